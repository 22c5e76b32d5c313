#!/usr/bin/env python
"""Device-resident timing of the other BASELINE configs (3 and 5) and of K sweeps on config 2 — parity-test cases in
the contract, measured here to see which roofline binds each.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _bootstrap  # noqa: E402

pkg = _bootstrap.load_package()
S = pkg.signals
HBM = 6514.8


def bytes_per_stream(k, block, parts):
    s1 = 8 * (block + 1)
    return 16 * block * k + 2 * s1 * ((parts - 1) + min(k, parts - 1)) + 16 * block + 320


def run(name, n_streams, block, taps, fs, decay, k_blocks, reps=5, eq=True):
    h = S.synthetic_hrir_set(taps, decay)
    eng = pkg.Engine(n_streams, block, taps, sample_rate=fs)
    eng.set_hrir_set(h)
    eng.eq_set_preset(S.EQ_PRESET_TYPICAL)
    eng.enable_timing()
    eng.set_eq_enable(eq)
    eng.set_gain(0.5)
    n = block * k_blocks
    x = torch.randn((n_streams, 2, n), device="cuda", dtype=torch.float32) * 0.1
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    parts = -(-taps // block)
    for _ in range(max(3, parts // k_blocks + 1)):
        eng.process_device(x.data_ptr(), y.data_ptr(), n)
    eng.sync()
    ms = []
    for _ in range(reps):
        eng.process_device(x.data_ptr(), y.data_ptr(), n)
        ms.append(eng.last_kernel_ms())
    t = float(np.median(ms)) * 1e-3
    audio = n_streams * n / fs
    algo = bytes_per_stream(k_blocks, block, parts) * n_streams
    print(json.dumps({"case": name, "n_streams": n_streams, "block": block, "taps": taps, "partitions": parts, "K": k_blocks,
                      "kernel_ms": t * 1e3, "stream_s_per_s": audio / t, "algorithmic_GBps": algo / t / 1e9,
                      "hbm_frac": algo / t / 1e9 / HBM, "in_out_MB": 2 * x.numel() * 4 / 1e6}))
    del eng, x, y
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]
    if "cfg2" in which:
        for k in (1, 8, 48, 192):
            run("cfg2", 1024, 256, 256, 48000.0, 40.0, k)
    if "cfg3" in which:
        for k in (1, 16, 128):
            run("cfg3 (one GPU's 8192 streams)", 8192, 128, 512, 48000.0, 80.0, k)
    if "cfg5" in which:
        for k in (1, 4, 8, 16, 32, 64, 128):
            run("cfg5 (one GPU's 256 streams)", 256, 1024, 48000, 96000.0, 0.15 * 96000.0, k)
