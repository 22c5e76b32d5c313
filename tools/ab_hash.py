#!/usr/bin/env python
"""Output hashes of a few renders (for bit-level A/B comparison of two builds: run once per OHS_LIB_OVERRIDE)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, numpy as np  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
out = {}
for name, (n_streams, block, taps, n_blocks, sets) in {"cfg2_like": (23, 256, 256, 12, 1), "two_sets": (9, 256, 200, 7, 2),
                                                        "b128": (10, 128, 128, 9, 1), "p2": (5, 256, 400, 6, 1)}.items():
    x = S.stream_inputs(n_streams, block * n_blocks, base_seed=77)
    e = pkg.Engine(n_streams, block, taps, n_hrir_sets=sets)
    for k in range(sets):
        e.set_hrir_set(S.synthetic_hrir_set(taps, 40.0, seed=5 + k), hrir_set=k)
    if sets > 1:
        for s in range(n_streams):
            e.bind_stream_hrir(s, s % sets)
    e.eq_set_preset(S.EQ_PRESET_TYPICAL); e.set_eq_enable(True); e.set_gain(0.5)
    y = e.process(x)
    out[name] = hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest()[:16]
print(json.dumps(out))
