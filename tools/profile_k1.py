#!/usr/bin/env python
"""ncu helper: K = 1 launches (the per-block API) of config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, torch  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
eng = pkg.Engine(1024, 256, 256); eng.set_hrir_set(S.synthetic_hrir_set(256, 40.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
eng.set_eq_enable(True); eng.set_gain(0.5)
n = 256 * 64
x = torch.randn((1024, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
for i in range(64):
    eng.process_device(x.data_ptr() + i * 1024, y.data_ptr() + i * 1024, 256, n)
eng.sync()
print("kernel ms", eng.last_kernel_ms())
