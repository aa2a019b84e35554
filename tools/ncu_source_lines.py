#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (read on stdin or a file):
which lines of the CUDA source issue the warp instructions of the profiled kernel.  Needs -lineinfo at compile time."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, samp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr is None or len(r) <= ie or r[2] != "-":      # only the per-source-line rows (address column "-")
        continue
    try:
        out.append((float(r[ie]), cur_file, r[0], r[1].strip()[:120], int(float(r[samp]))))
    except ValueError:
        pass
tot = sum(o[0] for o in out)
ts = sum(o[4] for o in out)
print(f"total warp instructions {tot:.0f}, stall samples {ts}")
for n, f, line, src, s in sorted(out, reverse=True)[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s / max(ts, 1) * 100:5.1f}% samp  {f}:{line:>4}  {src}")
