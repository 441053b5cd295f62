#!/usr/bin/env python
"""Top source lines by stall samples from `ncu --page source --csv --print-source cuda,sass`; usage: ncu_lines.py file.csv [N] [filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else ""
cur_file = ""
data = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] != "-": continue          # per-source-line summary rows have '-' in the SASS columns
    try:
        data.append((int(r[6] or 0), int(r[7] or 0), cur_file, r[0], r[1].strip()[:105]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}")
sel = [d for d in data if flt in d[2]]
for d in sorted(sel, key=lambda d: -d[0])[:N]:
    print(f"{100*d[0]/tot:5.2f}% ex={d[1]:11d} {d[2]}:{d[3]:>4s} {d[4]}")
