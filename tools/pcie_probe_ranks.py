#!/usr/bin/env python
"""Host-link ceiling with 1..N GPUs copying at once (the limiter of the end-to-end path at N > 1).
Launch under torchrun (one rank per GPU).  Phases, each bracketed by barriers: rank 0 alone, then the first 2, 4, ... N
ranks together; per phase H2D only, D2H only and both directions at once, 256 MiB pinned buffers, 6 repetitions.
Rank 0 prints one JSON line: per-phase per-rank GB/s (each way) and the aggregate.  Optionally pins each rank's thread
and pinned allocation to the CPUs next to its GPU when `nvidia-smi topo` lists an affinity (OHS_PROBE_PIN=1)."""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pin = os.environ.get("OHS_PROBE_PIN", "0") == "1"
affinity = None
if pin:
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(local)], capture_output=True, text=True).stdout
        # "CPU Affinity: 0-15" style line
        for line in out.splitlines():
            if "ffinity" in line and ":" in line:
                spec = line.split(":")[1].strip().split()[0]
                cpus = set()
                for part in spec.split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                os.sched_setaffinity(0, cpus)
                affinity = sorted(cpus)
                break
    except Exception:
        affinity = None
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()


def run(active, h2d, d2h, reps=6):
    torch.cuda.synchronize(); barrier()
    gbs = 0.0
    if active:
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        gbs = reps * n / (time.perf_counter() - t0) / 1e9
    barrier()
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    if world > 1:
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return [float(x.item()) for x in g]
    return [gbs]


for _ in range(2):
    run(True, True, True, 2)
phases = {}
k = 1
while k <= world:
    active = rank < k
    res = {}
    for name, a, b in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
        per_rank = run(active, a, b)[:k]
        res[name] = {"per_gpu_gbs_each_way": [round(x, 2) for x in per_rank], "aggregate_gbs_each_way": round(sum(per_rank), 1),
                     "min_gbs": round(min(per_rank), 2)}
    phases["%d_gpus" % k] = res
    k *= 2
if rank == 0:
    info = {"world": world, "host_cpus": os.cpu_count(), "pinned_threads": pin, "cpu_affinity_rank0": affinity, "bytes_per_copy": n, "phases": phases}
    try:
        info["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:4000]
    except Exception:
        pass
    print(json.dumps(info))
if world > 1:
    dist.destroy_process_group()
