"""Quick timings that isolate the one-band-per-lane EQ loop: config 5's time-batched call (spectra-only render pass),
config 5 block by block, and a one-stream EQ-only engine at block 1024 (config 4's bus EQ)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import _bootstrap
pkg = _bootstrap.load_package(); S = pkg.signals

def timed(eng, x, y, n, reps=5):
    eng.enable_timing()
    for _ in range(2): eng.process_device(x.data_ptr(), y.data_ptr(), n)
    eng.sync(); ms = []
    for _ in range(reps):
        eng.process_device(x.data_ptr(), y.data_ptr(), n); ms.append(eng.last_kernel_ms())
    return sorted(ms)[len(ms) // 2]

c = S.CONFIGS[5]
eng = pkg.Engine(256, 1024, c["taps"], sample_rate=c["fs"]); eng.set_hrir_set(S.synthetic_hrir_set(c["taps"], c["decay"]))
eng.eq_set_preset(S.EQ_PRESET_TYPICAL); eng.set_eq_enable(True); eng.set_gain(0.5)
n = 1024 * 64
x = torch.randn((256, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
ms = timed(eng, x, y, n)
print(json.dumps({"case": "cfg5 time-batched K=64", "ms": ms, "stream_s_per_s": 256 * n / c["fs"] / (ms * 1e-3)}))
del eng
post = pkg.Engine(1, 1024, 1, sample_rate=48000.0); post.set_conv_enable(False); post.eq_set_preset(S.EQ_PRESET_TYPICAL); post.set_eq_enable(True)
n = 96 * 256
x = torch.randn((1, 2, n), device="cuda") * 0.1; y = torch.empty_like(x); torch.cuda.synchronize()
ms = timed(post, x, y, n)
print(json.dumps({"case": "bus EQ, 1 stream, %d frames" % n, "ms": ms, "cycles_per_step": ms * 1e-3 * 1.965e9 / n}))
