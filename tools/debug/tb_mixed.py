import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import _bootstrap
ohs = _bootstrap.load_package()
S = ohs.signals
from oracle import oracle as O

block, taps, n_streams, n_blocks = 512, 5000, 3, 24
h = S.synthetic_hrir_set(taps, taps / 5.0, seed=21)
n = block * n_blocks
x = S.stream_inputs(n_streams, n, base_seed=1200)
coeffs = np.stack([ohs.eq_design(t, 48000.0, fc, q, g) for (t, fc, q, g) in S.EQ_PRESET_TYPICAL])
for eq in (True, False):
    if eq:
        ref, _ = O.render_batch(x, block, h, coeffs, [1] * 10, True, 0.7, n_threads=8)
    else:
        ref, _ = O.render_batch(x, block, h, np.zeros((1, 5), np.float32), [0], False, 0.7, n_threads=8)
    e = ohs.Engine(n_streams, block, taps)
    e.set_hrir_set(h); e.eq_set_preset(S.EQ_PRESET_TYPICAL); e.set_eq_enable(eq); e.set_gain(0.7)
    cut1 = 16 * block
    e.set_time_batch(True)
    parts = [e.process(x[:, :, :cut1])]
    e.set_time_batch(False)
    for b in range(16, n_blocks):
        parts.append(e.process(x[:, :, b * block:(b + 1) * block]))
    y = np.concatenate(parts, axis=2)
    err = np.abs(y - ref).reshape(n_streams, 2, n_blocks, block).max(axis=(0, 1, 3))
    print("eq", eq, "per-block max err", " ".join("%.1e" % v for v in err))
