"""Host-side rows next to the hot path: HRIR selection/wiring and AutoEQ CSV ingestion.  CPU only."""
import os

import numpy as np
import pytest

from open_headstage_b200 import autoeq, sofa


def test_nearest_neighbour_selection_matches_survey_indices(cipic):
    h = sofa.from_arrays(cipic["ir"], cipic["pos"], float(cipic["fs"]))
    assert h.filter_length == 200
    assert h.nearest(30.0, 0.0) == 308     # left speaker
    assert h.nearest(330.0, 0.0) == 908    # right speaker
    assert h.nearest(0.0, 0.0) == 608
    assert h.nearest(-30.0, 0.0) == 908    # azimuth wraps
    assert sofa.ui_azimuth_to_sofa(-30.0) == 30.0 and sofa.ui_azimuth_to_sofa(30.0) == 330.0
    l, r = h.get_hrtf_irs(30.0, 0.0)
    assert np.array_equal(l, cipic["ir"][308, 0]) and np.array_equal(r, cipic["ir"][308, 1])
    # every measurement is its own nearest neighbour: selection is an exact index lookup
    for i in range(0, 1250, 97):
        assert h.nearest(float(cipic["pos"][i, 0]), float(cipic["pos"][i, 1])) == i


def test_wire_speakers_calls_set_ir_for_the_four_paths(cipic):
    calls = []

    class FakeEngine:
        def set_ir(self, path, ir, hrir_set=0):
            calls.append((path, np.asarray(ir).copy(), hrir_set))

    h = sofa.from_arrays(cipic["ir"], cipic["pos"], 44100.0)
    il, ir_ = sofa.wire_speakers(FakeEngine(), h, sofa.ui_azimuth_to_sofa(-30.0), 0.0, sofa.ui_azimuth_to_sofa(30.0), 0.0)
    assert (il, ir_) == (308, 908)
    assert [c[0] for c in calls] == [0, 1, 2, 3]
    assert np.array_equal(calls[0][1], cipic["ir"][308, 0]) and np.array_equal(calls[3][1], cipic["ir"][908, 1])


@pytest.mark.skipif(not os.path.exists("/root/reference/data/hrtf/subject_003.sofa"), reason="reference tree only exists in the dev container")
def test_load_sofa_reads_the_bundled_file_without_hdf5(cipic):
    h = sofa.load_sofa("/root/reference/data/hrtf/subject_003.sofa")
    assert h.ir.shape == (1250, 2, 200)
    assert np.array_equal(h.ir, cipic["ir"]) and np.allclose(h.position, cipic["pos"])


def test_parse_autoeq_csv():
    text = "Filter-Type,Fc,Q,Gain\nLS,105,0.7,6.5\nPK,60,1.2,-3.0\nHS,10000,0.7,-4.0\n"
    bands = autoeq.parse_autoeq_csv(text)
    assert [(b.filter_type, b.frequency, b.q, b.gain, b.enabled) for b in bands] == [
        (autoeq.LOWSHELF, 105.0, 0.7, 6.5, True), (autoeq.PEAK, 60.0, 1.2, -3.0, True), (autoeq.HIGHSHELF, 10000.0, 0.7, -4.0, True)]
    with pytest.raises(ValueError):
        autoeq.parse_autoeq_csv("Filter-Type,Fc,Q,Gain\nLP,100,0.7,0\n")   # src/autoeq_parser.rs:48

    applied = []

    class FakeEngine:
        def eq_update_band(self, *a):
            applied.append(a)

    autoeq.apply_to_engine(FakeEngine(), bands)
    assert applied[0] == (0, autoeq.LOWSHELF, 105.0, 0.7, 6.5, True, 0) and len(applied) == 3
