#!/usr/bin/env python
"""ncu helper: run a few launches of the config-2 render kernel with the EQ and/or convolution stage switched off,
so that each warp role can be profiled without the other's contention.  usage: profile_stage.py <conv 0|1> <eq 0|1> [K]"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import _bootstrap, torch  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
conv, eq = bool(int(sys.argv[1])), bool(int(sys.argv[2]))
K = int(sys.argv[3]) if len(sys.argv) > 3 else 24
eng = pkg.Engine(1024, 256, 256); eng.set_hrir_set(S.synthetic_hrir_set(256, 40.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
eng.set_eq_enable(eq); eng.set_conv_enable(conv); eng.set_gain(0.5)
n = 256 * K
x = torch.randn((1024, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
for _ in range(5):
    eng.process_device(x.data_ptr(), y.data_ptr(), n)
eng.sync()
print("kernel ms", eng.last_kernel_ms())
