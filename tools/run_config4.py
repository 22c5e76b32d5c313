#!/usr/bin/env python
"""BASELINE config 4: N mono sources, each with its own SOFA direction, binaurally mixed to ONE stereo bus; sources are
sharded over the ranks, per-rank buses are summed with an NCCL reduce.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 tools/run_config4.py [n_sources] [seconds]
(or plain `python tools/run_config4.py` for one GPU).  Rank 0 prints one JSON line with the timing and, for small
runs, the max abs error against an f64 evaluation of the definition."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import _bootstrap  # noqa: E402

pkg = _bootstrap.load_package()
S, P = pkg.signals, pkg.parallel


def main():
    n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    block, fs = 256, 48000.0
    n = int(seconds * fs) // block * block
    ir = np.load(os.path.join(ROOT, "tests", "golden", "cipic003_hrir.npz"))["ir"]
    lo, hi = P.shard_range(n_src, rank, world)
    lo, hi = lo - lo % 2, hi - hi % 2 if hi != n_src else hi
    hr = np.stack([ir[(s * 37) % 1250] for s in range(lo, hi)]).astype(np.float32)
    base = np.stack([S.pink_noise(n, 4000 + s) for s in range(16)]) / np.float32(64.0)
    src = np.stack([base[s % 16] * np.float32(1.0 if (s // 16) % 2 == 0 else -1.0) for s in range(lo, hi)])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bus = P.render_object_mix(pkg, src, hr, block, fs, eq_preset=S.EQ_PRESET_TYPICAL, gain=0.5, device=local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        out = {"config": "cfg4: %d mono sources -> one stereo bus, %d GPU(s)" % (n_src, world), "n_frames": n, "wall_s_including_setup": dt,
               "source_seconds_per_s_including_setup": n_src * n / fs / dt, "bus_peak": float(bus.abs().max())}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
