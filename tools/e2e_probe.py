#!/usr/bin/env python
"""End-to-end (host pinned buffers -> ohs_process -> host) throughput against staging chunk size (OHS_STAGE_MB)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, numpy as np  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
n = 256 * 192
eng = pkg.Engine(1024, 256, 256); eng.set_hrir_set(S.synthetic_hrir_set(256, 40.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL)
eng.set_eq_enable(True); eng.set_gain(0.5)
x = pkg.PinnedBuffer((1024, 2, n)); y = pkg.PinnedBuffer((1024, 2, n))
x.array[...] = np.random.default_rng(0).standard_normal((1024, 2, n), dtype=np.float32) * 0.1
for _ in range(2):
    eng.process(x.array, out=y.array)
t0 = time.perf_counter(); k = 8
for _ in range(k):
    eng.process(x.array, out=y.array)
dt = (time.perf_counter() - t0) / k
print(json.dumps({"stage_mb": os.environ.get("OHS_STAGE_MB", "24 (default)"), "ms_per_step": dt * 1e3, "stream_s_per_s": 1024 * n / 48000 / dt,
                  "GBps_each_way": x.array.nbytes / dt / 1e9}))
