#!/usr/bin/env python
"""How fast is each stage when an SM has plenty of streams?  Config-2 geometry (block 256, 256 taps) with 8192 streams
(55 per SM) and the EQ and/or convolution stage switched off; prints streams per kilo-cycle per SM."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, numpy as np, torch  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
n_streams, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 24
def run(conv, eq):
    eng = pkg.Engine(n_streams, 256, 256); eng.set_hrir_set(S.synthetic_hrir_set(256, 40.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
    eng.set_eq_enable(eq); eng.set_conv_enable(conv); eng.set_gain(0.5)
    n = 256 * K
    x = torch.randn((n_streams, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
    for _ in range(3): eng.process_device(x.data_ptr(), y.data_ptr(), n)
    eng.sync(); ms = []
    for _ in range(5):
        eng.process_device(x.data_ptr(), y.data_ptr(), n); ms.append(eng.last_kernel_ms())
    t = float(np.median(ms)) * 1e-3
    cyc_per_block = t / K * 1.965e9
    print(json.dumps({"n_streams": n_streams, "conv": conv, "eq": eq, "ms": t * 1e3, "stream_s_per_s": n_streams * n / 48000 / t,
                      "streams_per_kcycle_per_sm": n_streams / 148 / (cyc_per_block / 1e3)}))
run(True, True); run(False, True); run(True, False)
