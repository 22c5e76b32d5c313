#!/usr/bin/env python
"""ncu helper: a few K-block launches of config 5 (256 streams, B 1024, 48 000 taps) through the time-batched path."""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import _bootstrap, torch  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = pkg.Engine(256, 1024, 48000, sample_rate=96000.0); eng.set_hrir_set(S.synthetic_hrir_set(48000, 9000.0)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
eng.set_eq_enable(True); eng.set_gain(0.5)
n = 1024 * K
x = torch.randn((256, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
for _ in range(3):
    eng.process_device(x.data_ptr(), y.data_ptr(), n)
eng.sync()
print("ms per launch", eng.last_kernel_ms())
