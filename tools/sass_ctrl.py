#!/usr/bin/env python3
"""Decode the scheduling control fields of `cuobjdump -sass` output (128-bit encodings, Volta-and-later layout):
stall count, yield, write/read scoreboard index, scoreboard wait mask.  Usage: sass_ctrl.py file.sass [lo_addr hi_addr]"""
import re, sys

def parse(path):
    rows, cur = [], None
    for line in open(path):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            rows.append(cur)
            continue
        m = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and cur is not None and cur[3] is None:
            cur[3] = int(m.group(1), 16)
    return rows

def ctrl(hi):
    return dict(stall=(hi >> 41) & 0xf, yld=(hi >> 45) & 1, wr=(hi >> 46) & 7, rd=(hi >> 49) & 7,
                wait=(hi >> 52) & 0x3f, reuse=(hi >> 58) & 0xf)

if __name__ == "__main__":
    rows = parse(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
    for addr, text, _, h in rows:
        if not (lo <= addr <= hi) or h is None:
            continue
        c = ctrl(h)
        wait = "".join(str(i) for i in range(6) if c["wait"] >> i & 1) or "-"
        print("%05x  st%-2d %s w%s r%s wait[%-6s]  %s" % (addr, c["stall"], "Y" if c["yld"] else " ",
              c["wr"] if c["wr"] != 7 else "-", c["rd"] if c["rd"] != 7 else "-", wait, text))
