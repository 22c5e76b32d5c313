"""Stage isolation: the render kernel with one role switched off (EQ only / convolution only / both / neither), for
config 2 (default) or config 3 (`cfg3`).  Prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _bootstrap  # noqa: E402

pkg = _bootstrap.load_package(); S = pkg.signals
CFG = {"cfg2": (1024, 256, 256, 40.0, 192), "cfg3": (8192, 128, 512, 80.0, 128)}


def run(cfg, conv, eq):
    n_streams, block, taps, decay, K = CFG[cfg]
    eng = pkg.Engine(n_streams, block, taps); eng.set_hrir_set(S.synthetic_hrir_set(taps, decay)); eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.enable_timing()
    eng.set_eq_enable(eq); eng.set_conv_enable(conv); eng.set_gain(0.5)
    n = block * K
    x = torch.randn((n_streams, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
    for _ in range(3):
        eng.process_device(x.data_ptr(), y.data_ptr(), n)
    eng.sync(); ms = []
    for _ in range(5):
        eng.process_device(x.data_ptr(), y.data_ptr(), n); ms.append(eng.last_kernel_ms())
    t = float(np.median(ms))
    print(json.dumps({"config": cfg, "conv": conv, "eq": eq, "ms": t, "us_per_block": t * 1e3 / K, "cycles_per_block": t * 1e-3 / K * 1.965e9,
                      "streams_per_cta": eng.streams_per_cta()}))


cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
run(cfg, True, True); run(cfg, False, True); run(cfg, True, False); run(cfg, False, False)
