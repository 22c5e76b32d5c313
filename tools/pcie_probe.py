#!/usr/bin/env python
"""PCIe ceiling on this box: pinned-host <-> device copies alone and both directions at once (what ohs_process's
3-stage pipeline can at best sustain)."""
import json, time, torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8, chunk=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
for _ in range(2): run(True, True, 2)
print(json.dumps({"h2d_only_GBps": run(True, False), "d2h_only_GBps": run(False, True), "both_GBps_each_way": run(True, True),
                  "both_24MiB_chunks": run(True, True, chunk=24 << 20), "both_4MiB_chunks": run(True, True, chunk=4 << 20)}))
