#!/usr/bin/env python
"""Per-instruction sample listing of one code region of an .ncu-rep: all instructions whose execution count equals EXEC
(the grouping `ncu_summary.py` prints).  Usage: ncu_region.py file.ncu-rep EXEC"""
import csv, subprocess, sys
rep, want = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
for r in data:
    if f(r, "Instructions Executed") != want: continue
    st = sorted(((k.replace("stall_", ""), int(f(r, k))) for k in keys), key=lambda kv: -kv[1])
    print("%s %5d  %-52s %s" % (r[0][-5:], f(r, "# Samples"), r[1][:52], [kv for kv in st if kv[1] and kv[0] != "selected"][:3]))
