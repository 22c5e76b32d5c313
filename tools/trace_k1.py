"""Where a K = 1 launch (one engine block per ohs_process_device call, the reference's calling pattern) spends its time.
Needs the instrumented library:  python tools/ab_build.py trace -DOHS_TRACE ;  OHS_LIB_OVERRIDE=.../libohs_cuda_trace.so
Prints, per milestone, the median over CTAs and launches of clock64 cycles since the CTA's first instruction, and the
CUDA-event time per launch with and without programmatic dependent launches.
usage: trace_k1.py [cfg=2] [blocks_per_launch=1] [launches=32]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import _bootstrap

pkg = _bootstrap.load_package()
S = pkg.signals
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
c = S.CONFIGS[cfg]
n_streams, block, taps, fs = c["n_streams"], c["block"], c["taps"], c["fs"]
NAMES = ["entry", "prologue done", "eq: coeffs loaded", "eq: first rows landed", "eq: last block filtered", "conv: first block ready",
         "conv: last block written", "conv: history saved", "eq: state saved", "stager: first copies issued", "stager: mbarriers initialised",
         "stager: constant tables issued", "stager: previous launch complete", "conv: block 0 and history rows issued"]


def run(pdl: str):
    os.environ["OHS_PDL"] = pdl
    e = pkg.Engine(n_streams, block, taps, sample_rate=fs)
    e.set_hrir_set(S.synthetic_hrir_set(taps, c["decay"]))
    e.eq_set_preset(S.EQ_PRESET_TYPICAL); e.set_eq_enable(True); e.set_gain(0.5)
    n = block * K * L
    x = torch.from_numpy(S.stream_inputs(n_streams, n, unique=8)).cuda()
    y = torch.empty_like(x)
    stream = torch.cuda.ExternalStream(e.cuda_stream())
    g = e.streams_per_cta()
    n_cta = (n_streams + g - 1) // g
    stamps = torch.zeros((L, n_cta, 16), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for i in range(8):
        e.process_device(x.data_ptr() + i * block * K * 4, y.data_ptr() + i * block * K * 4, block * K, n)
    e.sync()
    traced = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(L):
        try:
            e.debug_trace(stamps[i].data_ptr())
        except pkg.OhsError:
            traced = False
        e.process_device(x.data_ptr() + i * block * K * 4, y.data_ptr() + i * block * K * 4, block * K, n)
    e1.record(stream)
    e.sync()
    us = e0.elapsed_time(e1) * 1e3 / L
    out = {"cfg": cfg, "pdl": pdl, "blocks_per_launch": K, "streams_per_cta": g, "ctas": n_cta, "us_per_launch": us,
           "stream_s_per_s": n_streams * (block * K / fs) / (us * 1e-6)}
    if traced:
        st = stamps.cpu().numpy()[4:]          # skip the first launches
        rel = st - st[:, :, :1]
        out["cycles_since_entry_median"] = {NAMES[i]: float(np.median(rel[:, :, i][st[:, :, i] > 0])) for i in range(1, len(NAMES)) if (st[:, :, i] > 0).any()}
        out["cycles_since_entry_max"] = {NAMES[i]: float(np.max(rel[:, :, i][st[:, :, i] > 0])) for i in range(1, len(NAMES)) if (st[:, :, i] > 0).any()}
    print(json.dumps(out))


run("1")
run("0")
