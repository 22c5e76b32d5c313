#!/usr/bin/env python
"""Feasibility probe for the north star's tensor-core option: config 2's convolution as ONE block-Toeplitz GEMM
Y[(stream, block), 2 ears x B] = X[(stream, block), 2 channels x 2B window] . T[2 channels x 2B, 2 ears x B]
(T holds the four 256-tap HRIR paths as Toeplitz blocks, shared by all streams).  Library GEMM (torch.matmul = cuBLAS)
as an upper bound for a hand-written tcgen05 kernel: TF32 in one pass (accuracy check against f64) and the three-pass
hi/lo split that the 1e-5 parity bar needs.  Prints stream-s/s equivalents next to the FFT path's numbers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch  # noqa: E402
import _bootstrap  # noqa: E402
pkg = _bootstrap.load_package(); S = pkg.signals
B, TAPS, STREAMS, K = 256, 256, 1024, 192
h = S.synthetic_hrir_set(TAPS, 40.0).astype(np.float64)            # [4][taps]: LSL, LSR, RSL, RSR
T = np.zeros((2, 2 * B, 2, B))                                     # [channel][window sample][ear][out sample]
for ch in range(2):
    for ear in range(2):
        hp = h[2 * ch + ear]
        for n in range(B):                                         # y[n] = sum_k h[k] x[B + n - k]
            for k in range(TAPS):
                T[ch, B + n - k, ear, n] = hp[k]
T = T.reshape(4 * B, 2 * B)
x = S.stream_inputs(8, B * 9, base_seed=5).astype(np.float64)      # accuracy sample: 8 streams, 8 blocks with history
rows = np.stack([np.concatenate([x[s, 0, (b - 1) * B:(b + 1) * B], x[s, 1, (b - 1) * B:(b + 1) * B]]) for s in range(8) for b in range(1, 9)])
ref = rows @ T
dev = torch.device("cuda")
Tt = torch.tensor(T, dtype=torch.float32, device=dev); Rt = torch.tensor(rows, dtype=torch.float32, device=dev)
def tf32_mm(a, b):
    torch.backends.cuda.matmul.allow_tf32 = True
    return a @ b
def split(a):  # hi = a rounded to TF32's 10-bit mantissa, lo = the rest
    hi = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, a - hi
def mm3(a, b):
    ah, al = split(a); bh, bl = split(b)
    return tf32_mm(ah, bh) + (tf32_mm(ah, bl) + tf32_mm(al, bh))
torch.backends.cuda.matmul.allow_tf32 = False
e_fp32 = float(np.max(np.abs((Rt @ Tt).double().cpu().numpy() - ref)))
e_tf32 = float(np.max(np.abs(tf32_mm(Rt, Tt).double().cpu().numpy() - ref)))
e_3x = float(np.max(np.abs(mm3(Rt, Tt).double().cpu().numpy() - ref)))
# throughput: one bench step = 1024 streams x 192 blocks rows
M = STREAMS * K
A = torch.randn((M, 4 * B), device=dev) * 0.1
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
ms1 = timeit(lambda: tf32_mm(A, Tt))
Ah, Al = split(A); Th, Tl = split(Tt)
ms3 = timeit(lambda: (tf32_mm(Ah, Th), tf32_mm(Ah, Tl), tf32_mm(Al, Th)))
sec = STREAMS * K * B / 48000.0
flop = 2.0 * M * 4 * B * 2 * B
print(json.dumps({"gemm": "M=%d K=%d N=%d" % (M, 4 * B, 2 * B), "max_abs_err_fp32_gemm": e_fp32, "max_abs_err_tf32_1pass": e_tf32,
                  "max_abs_err_tf32_3pass": e_3x, "tf32_1pass_ms": ms1, "tf32_1pass_tflops": flop / ms1 / 1e9,
                  "tf32_1pass_stream_s_per_s": sec / (ms1 * 1e-3), "tf32_3pass_ms": ms3, "tf32_3pass_stream_s_per_s": sec / (ms3 * 1e-3),
                  "note": "convolution only: no EQ, no window gather, no hi/lo split cost, inputs resident; FFT path (EQ off): 9.84k cycles/block = 1.09 M stream-s/s, whole chain 1.077 M"}))
