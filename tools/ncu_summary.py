#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall mix per code region (grouped by
execution count, which separates the EQ warps' loops from the convolution warps' passes), top stall sites and shared
memory bank conflicts.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 20]"""
import collections
import csv
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    det = run([rep, "--page", "details"])
    keys = ("Duration", "Elapsed Cycles", "SM Frequency", "Registers Per", "Achieved Occ", "Theoretical Occ", "Executed Ipc",
            "Issue Slots Busy", "Block Limit", "Waves Per", "bank conflicts", "DRAM Throughput", "Dynamic Shared", "Grid Size",
            "Block Size", "L1/TEX Hit", "L2 Hit")
    for line in det.splitlines():
        if any(k in line for k in keys):
            print(line.rstrip()[:150])
    raw = list(csv.reader(run([rep, "--page", "raw", "--csv"]).splitlines()))
    if len(raw) >= 3:
        for h, u, v in zip(raw[0], raw[1], raw[2]):
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__inst_executed.max",
                     "smsp__inst_executed.min", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                     "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                     "sm__cycles_active.min", "sm__cycles_active.max", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"):
                print("%-70s %-8s %s" % (h, u, v))
    rows = list(csv.reader(run([rep, "--page", "source", "--csv"]).splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(f(r, "# Samples") for r in data)
    print("\ntotal samples %d, instructions %d, warp-instr executed %.3e" % (tot, len(data), sum(f(r, "Instructions Executed") for r in data)))
    groups = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
    for r in data:
        g = groups[f(r, "Instructions Executed")]
        g[0] += 1
        g[1] += f(r, "# Samples")
        for k in stall_keys:
            g[2][k.replace("stall_", "")] += f(r, k)
    print("regions by execution count:")
    for e, (n, sm, c) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:10]:
        print("  exec=%10.0f n_instr=%4d samples=%6.0f (%4.1f%%) %s" % (e, n, sm, 100 * sm / max(tot, 1), [(k, int(v)) for k, v in c.most_common(5)]))
    print("top stall sites:")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
        st = sorted(((k.replace("stall_", ""), f(r, k)) for k in stall_keys), key=lambda kv: -kv[1])[:2]
        print("  %s %6.0f %5.1f%% exec=%9.0f %-58s %s" % (r[0][-5:], f(r, "# Samples"), 100 * f(r, "# Samples") / max(tot, 1),
                                                      f(r, "Instructions Executed"), r[1][:58], [(k, int(v)) for k, v in st if v]))
    print("shared-memory excessive wavefronts:")
    for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:8]:
        if f(r, "L1 Wavefronts Shared Excessive") > 0:
            print("  %s %-50s exc=%d total=%d ideal=%d" % (r[0][-5:], r[1][:50], f(r, "L1 Wavefronts Shared Excessive"),
                                                          f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal")))


if __name__ == "__main__":
    main()
