#!/usr/bin/env python
"""Small multi-block render (7 streams per CTA path, P > 1, EQ on) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, numpy as np  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
os.environ.setdefault("OHS_STREAMS_PER_CTA", "7")
for (n_streams, block, taps) in ((9, 256, 600), (5, 128, 128)):
    eng = pkg.Engine(n_streams, block, taps)
    eng.set_hrir_set(S.synthetic_hrir_set(taps, 60.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.set_eq_enable(True); eng.set_gain(0.5)
    x = S.stream_inputs(n_streams, block * 5)
    y = eng.process(x)
    y2 = eng.process(x[:, :, :block])
    print("ok", n_streams, block, float(np.abs(y).max()), float(np.abs(y2).max()))
# the time-batched route: EQ pre-pass (thin shape and the SM-owning shape), forward, per-bin convolution, inverse,
# several chunks per call, a second call that starts from the first one's state
os.environ.pop("OHS_STREAMS_PER_CTA", None)
for (n_streams, block, taps, n_blocks) in ((3, 128, 1100, 40), (14, 64, 600, 72)):
    eng = pkg.Engine(n_streams, block, taps)
    eng.set_hrir_set(S.synthetic_hrir_set(taps, 200.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.set_eq_enable(True); eng.set_gain(0.5)
    x = S.stream_inputs(n_streams, block * n_blocks)
    y = eng.process(x)
    y2 = eng.process(x[:, :, :block * 9])
    print("ok time-batched", n_streams, block, float(np.abs(y).max()), float(np.abs(y2).max()))
