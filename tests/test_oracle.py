"""Pins the CPU oracle (oracle/ohs_oracle.c) before anything is compared with it.

1. The reference's own five DSP unit tests, replayed verbatim (src/dsp/convolution.rs:317-421,
   src/dsp/parametric_eq.rs:218-238) — the only known-answer vectors the reference holds for the hot path.
2. Independent f64 evaluations (numpy/scipy) of the same mathematics: direct convolution, scipy.signal.lfilter with
   the same f32 coefficients, RBJ cookbook formulae in f64.
"""
import numpy as np
import pytest
import scipy.signal as sps

from oracle import oracle as O
from open_headstage_b200 import signals as S

BLOCK_SIZE = 512  # src/dsp/convolution.rs:22
TOLERANCE = 1e-3  # src/dsp/convolution.rs:301


def test_ref_identity_ir_passthrough():
    """src/dsp/convolution.rs:317-347"""
    e = O.ConvolutionEngine(BLOCK_SIZE)
    e.set_ir(O.LSL, [1.0]); e.set_ir(O.LSR, [0.0]); e.set_ir(O.RSL, [0.0]); e.set_ir(O.RSR, [1.0])
    i = np.arange(BLOCK_SIZE, dtype=np.float32)
    in_l = np.sin(i * np.float32(0.1)).astype(np.float32)
    in_r = np.sin(i * np.float32(-0.1)).astype(np.float32)
    e.process_block(in_l, in_r)
    out_l, out_r = e.process_block(in_l, in_r)
    assert np.max(np.abs(out_l - in_l)) < TOLERANCE
    assert np.max(np.abs(out_r - in_r)) < TOLERANCE


def test_ref_delay_ir():
    """src/dsp/convolution.rs:349-383"""
    e = O.ConvolutionEngine(BLOCK_SIZE)
    d = 5
    ir = np.zeros(d + 1, np.float32); ir[d] = 1.0
    e.set_ir(O.LSL, ir); e.set_ir(O.LSR, [0.0]); e.set_ir(O.RSL, [0.0]); e.set_ir(O.RSR, [0.0])
    in_l = np.arange(BLOCK_SIZE * 2, dtype=np.float32)
    in_r = np.zeros(BLOCK_SIZE * 2, np.float32)
    out_l, _ = e.process_block(in_l, in_r)
    expected = np.zeros_like(in_l); expected[d:] = in_l[:-d]
    assert np.max(np.abs(out_l[d:] - expected[d:])) < TOLERANCE


def test_ref_long_ir_partitioning():
    """src/dsp/convolution.rs:385-421"""
    e = O.ConvolutionEngine(BLOCK_SIZE)
    ir_len = BLOCK_SIZE + BLOCK_SIZE // 2
    ir = np.zeros(ir_len, np.float32); ir[0] = 1.0; ir[-1] = 0.5
    e.set_ir(O.LSL, ir)
    assert e.num_partitions(O.LSL) == 2
    in_l = np.zeros(BLOCK_SIZE * 3, np.float32); in_l[0] = 1.0
    out_l, _ = e.process_block(in_l, np.zeros_like(in_l))
    expected = np.zeros_like(in_l); expected[0] = 1.0; expected[ir_len - 1] = 0.5
    assert np.max(np.abs(out_l[:ir_len] - expected[:ir_len])) < TOLERANCE


def test_ref_biquad_passthrough_when_disabled():
    """src/dsp/parametric_eq.rs:218-225 — exact equality"""
    q = O.StereoParametricEQ(1, 48000.0)
    l, r = q.process_block([0.5], [0.5])
    assert l[0] == np.float32(0.5) and r[0] == np.float32(0.5)


def test_ref_biquad_processes_when_enabled():
    """src/dsp/parametric_eq.rs:227-238"""
    q = O.StereoParametricEQ(1, 48000.0)
    q.update_band_coeffs(0, 48000.0, O.LOWPASS, 1000.0, 0.707, 0.0, True)
    l, _ = q.process_block([0.5], [0.5])
    assert l[0] != np.float32(0.5)


def test_default_engine_is_silent_and_empty_ir_mutes():
    """ConvolutionPathData::new default IR = zeros (:46-48); empty slice -> one silent partition (:114-118)"""
    e = O.ConvolutionEngine(128)
    x = S.pink_noise(256, 3)
    l, r = e.process_block(x, x)
    assert not l.any() and not r.any()
    e.set_ir(O.LSL, [1.0])
    e.set_ir(O.LSL, [])
    assert e.num_partitions(O.LSL) == 1
    l, r = e.process_block(x, x)
    assert not l.any() and not r.any()


@pytest.mark.parametrize("n,expect_zero", [(512, False), (256, True), (100, True), (1024, False)])
def test_fifo_latency_semantics(n, expect_zero):
    """process_block FIFO adaptation (:141-182): host block not a multiple of 512 -> the first call is zero-filled."""
    e = O.ConvolutionEngine(BLOCK_SIZE)
    e.set_ir(O.LSL, [1.0]); e.set_ir(O.RSR, [1.0])
    x = S.pink_noise(n, 5)
    l, _ = e.process_block(x, x)
    assert (not l.any()) == expect_zero
    if n == 256:  # second call completes the 512 block: output = first 256 inputs (latency 512 - n)
        l2, _ = e.process_block(x, x)
        assert np.max(np.abs(l2 - x)) < 1e-6


@pytest.mark.parametrize("block,taps", [(64, 200), (128, 512), (256, 256), (512, 200), (1024, 5000)])
def test_conv_matches_f64_direct_convolution(block, taps):
    h = S.synthetic_hrir_set(taps, taps / 6.0, seed=11)
    n = block * 12
    xl, xr = S.pink_noise(n, 21), S.pink_noise(n, 22)
    e = O.ConvolutionEngine(block)
    for p in range(4):
        assert e.set_ir(p, h[p]) == -(-taps // block)
    out_l, out_r = e.process_block(xl, xr)
    h64, l64, r64 = h.astype(np.float64), xl.astype(np.float64), xr.astype(np.float64)
    tl = (np.convolve(l64, h64[O.LSL]) + np.convolve(r64, h64[O.RSL]))[:n]  # :229
    tr = (np.convolve(l64, h64[O.LSR]) + np.convolve(r64, h64[O.RSR]))[:n]  # :230
    assert np.max(np.abs(out_l - tl)) < 2e-6
    assert np.max(np.abs(out_r - tr)) < 2e-6


def _rbj_f64(t, fs, fc, q, g):
    w = 2 * np.pi * fc / fs
    s, c = np.sin(w), np.cos(w)
    al = s / (2 * q)
    a = 10 ** (g / 40)
    sq = 2 * al * np.sqrt(a)
    if t == O.PEAK:
        b = [1 + al * a, -2 * c, 1 - al * a]; d = [1 + al / a, -2 * c, 1 - al / a]
    elif t == O.LOWSHELF:
        b = [a * ((a + 1) - (a - 1) * c + sq), 2 * a * ((a - 1) - (a + 1) * c), a * ((a + 1) - (a - 1) * c - sq)]
        d = [(a + 1) + (a - 1) * c + sq, -2 * ((a - 1) + (a + 1) * c), (a + 1) + (a - 1) * c - sq]
    elif t == O.HIGHSHELF:
        b = [a * ((a + 1) + (a - 1) * c + sq), -2 * a * ((a - 1) + (a + 1) * c), a * ((a + 1) + (a - 1) * c - sq)]
        d = [(a + 1) - (a - 1) * c + sq, 2 * ((a - 1) - (a + 1) * c), (a + 1) - (a - 1) * c - sq]
    elif t == O.LOWPASS:
        b = [(1 - c) / 2, 1 - c, (1 - c) / 2]; d = [1 + al, -2 * c, 1 - al]
    elif t == O.HIGHPASS:
        b = [(1 + c) / 2, -(1 + c), (1 + c) / 2]; d = [1 + al, -2 * c, 1 - al]
    elif t == O.BANDPASS:
        b = [s / 2, 0, -s / 2]; d = [1 + al, -2 * c, 1 - al]
    elif t == O.NOTCH:
        b = [1, -2 * c, 1]; d = [1 + al, -2 * c, 1 - al]
    else:
        b = [1 - al, -2 * c, 1 + al]; d = [1 + al, -2 * c, 1 - al]
    return np.array([b[0], b[1], b[2], d[1], d[2]]) / d[0]


@pytest.mark.parametrize("t", range(8))
def test_eq_design_matches_rbj_cookbook_f64(t):
    for fs, fc, q, g in [(48000.0, 1000.0, 0.707, 6.0), (96000.0, 105.0, 0.7, -4.5), (44100.0, 9800.0, 2.5, 3.0)]:
        c = O.eq_design(t, fs, fc, q, g)
        ref = _rbj_f64(t, fs, fc, q, g)
        assert np.allclose(c, ref, rtol=2e-5, atol=2e-6), (t, c, ref)


def test_eq_design_error_codes():
    """from_params rejects 2*f0 > fs and q < 0 (the reference .unwrap()s, src/dsp/parametric_eq.rs:111)"""
    with pytest.raises(ValueError):
        O.eq_design(O.PEAK, 48000.0, 30000.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        O.eq_design(O.PEAK, 48000.0, 1000.0, -1.0, 0.0)


def test_eq_cascade_matches_f64_lfilter_with_same_coefficients():
    fs, n = 48000.0, 48000
    q = O.StereoParametricEQ(10, fs)
    coeffs = []
    for i, (t, fc, qq, g) in enumerate(S.EQ_PRESET_TYPICAL):
        q.update_band_coeffs(i, fs, t, fc, qq, g, True)
        coeffs.append(O.eq_design(t, fs, fc, qq, g).astype(np.float64))
    xl, xr = S.pink_noise(n, 1), S.pink_noise(n, 2)
    l, r = q.process_block(xl, xr)
    tl, tr = xl.astype(np.float64), xr.astype(np.float64)
    for c in coeffs:
        tl = sps.lfilter(c[:3], [1.0, c[3], c[4]], tl)
        tr = sps.lfilter(c[:3], [1.0, c[3], c[4]], tr)
    # the f32 cascade carries ~2e-4 of its own round-off on this preset (SURVEY.md §8d probes)
    assert np.max(np.abs(l - tl)) < 2e-3
    assert np.max(np.abs(r - tr)) < 2e-3
    # block-split invariance: state carries across calls exactly
    q2 = O.StereoParametricEQ(10, fs)
    for i, (t, fc, qq, g) in enumerate(S.EQ_PRESET_TYPICAL):
        q2.update_band_coeffs(i, fs, t, fc, qq, g, True)
    parts = [q2.process_block(xl[a:a + 777], xr[a:a + 777]) for a in range(0, n, 777)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), l)
    q2.reset_all_bands_state()
    assert not q2.state().any()


def test_frequency_response_matches_freqz():
    fs = 48000.0
    q = O.StereoParametricEQ(10, fs)
    h = np.ones(64, np.complex128)
    freqs = np.geomspace(20, 20000, 64)
    for i, (t, fc, qq, g) in enumerate(S.EQ_PRESET_TYPICAL):
        en = i != 3
        q.update_band_coeffs(i, fs, t, fc, qq, g, en)
        if en:
            c = O.eq_design(t, fs, fc, qq, g).astype(np.float64)
            h *= sps.freqz(c[:3], [1.0, c[3], c[4]], worN=freqs, fs=fs)[1]
    got = q.calculate_frequency_response(fs, freqs)
    assert np.allclose(got, np.abs(h), rtol=2e-3)


def test_chain_order_bypass_gain(cipic):
    """src/lib.rs:1169-1207: bypass leaves the buffer untouched; order is EQ -> conv -> gain."""
    ir = cipic["ir"]
    e = O.ConvolutionEngine(512)
    e.set_ir(O.LSL, ir[308, 0]); e.set_ir(O.LSR, ir[308, 1]); e.set_ir(O.RSL, ir[908, 0]); e.set_ir(O.RSR, ir[908, 1])
    q = O.StereoParametricEQ(10, 48000.0)
    for i, (t, fc, qq, g) in enumerate(S.EQ_PRESET_TYPICAL):
        q.update_band_coeffs(i, 48000.0, t, fc, qq, g, True)
    xl, xr = S.pink_noise(2048, 1), S.pink_noise(2048, 2)
    bl, br = O.chain_process(e, q, True, True, 0.5, xl, xr)
    assert np.array_equal(bl, xl) and np.array_equal(br, xr)
    yl, yr = O.chain_process(e, q, True, False, 0.5, xl, xr)
    # recompute by hand
    e2 = O.ConvolutionEngine(512)
    e2.set_ir(O.LSL, ir[308, 0]); e2.set_ir(O.LSR, ir[308, 1]); e2.set_ir(O.RSL, ir[908, 0]); e2.set_ir(O.RSR, ir[908, 1])
    q2 = O.StereoParametricEQ(10, 48000.0)
    for i, (t, fc, qq, g) in enumerate(S.EQ_PRESET_TYPICAL):
        q2.update_band_coeffs(i, 48000.0, t, fc, qq, g, True)
    el, er = q2.process_block(xl, xr)
    cl, cr = e2.process_block(el, er)
    assert np.array_equal(yl, cl * np.float32(0.5)) and np.array_equal(yr, cr * np.float32(0.5))


def test_render_batch_threads_agree():
    h = S.synthetic_hrir_set(256, 40.0)
    bc = np.stack([O.eq_design(t, 48000.0, fc, q, g) for (t, fc, q, g) in S.EQ_PRESET_TYPICAL])
    x = S.stream_inputs(5, 1024)
    y1, _ = O.render_batch(x, 256, h, bc, [1] * 10, True, 0.5, n_threads=1)
    y3, _ = O.render_batch(x, 256, h, bc, [1] * 10, True, 0.5, n_threads=3)
    assert np.array_equal(y1, y3)
    assert np.abs(y1).max() > 0.01
