#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) by enclosing function symbol ranges; usage: ncu_hot.py src.csv lib.so"""
import csv, sys, subprocess, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isamp, iexec, inoinst, ilong, iwait = [hdr.index(x) for x in ("Address", "# Samples", "Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait")]
isrc = hdr.index("Source")
# symbol ranges from the ELF (offsets inside the kernel's .text)
elf = subprocess.run(["cuobjdump", "-elf", sys.argv[2]], capture_output=True, text=True).stdout
kname = sys.argv[3] if len(sys.argv) > 3 else "optimize_kernelILi2ELi10ELi32"
shndx = None
for m in re.finditer(r"^\s*0x[0-9a-f]+\s+(?:0x[0-9a-f]+|0)\s+(?:0x[0-9a-f]+|0)\s+0x12\s+\S+\s+(0x[0-9a-f]+)\s+(\S+)", elf, re.M):
    if kname in m.group(2): shndx = m.group(1)
syms = []
for m in re.finditer(r"^\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+)\s+0x(?:2|22)\s+\S+\s+(0x[0-9a-f]+)\s+(\S+)", elf, re.M):
    if m.group(3) == shndx: syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
syms.sort()
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows[2:]:
    if len(r) <= isamp or not r[ia]: continue
    a = int(r[ia], 16)
    if base is None: base = a
    off = a - base
    name = "kernel_main"
    for s0, sz, nm in syms:
        if s0 <= off < s0 + sz: name = nm.split("$")[-1]
    f = lambda i: int(r[i] or 0)
    g = agg[name]
    g[0] += f(isamp); g[1] += f(iexec); g[2] += f(inoinst); g[3] += f(ilong); g[4] += f(iwait)
tot = sum(v[0] for v in agg.values()) or 1
print(f"{'function':60s} {'samples%':>8s} {'inst_exec':>12s} {'no_inst%':>8s} {'long_sb%':>8s} {'wait%':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:60]:60s} {100*v[0]/tot:8.1f} {v[1]:12d} {100*v[2]/max(1,v[0]):8.1f} {100*v[3]/max(1,v[0]):8.1f} {100*v[4]/max(1,v[0]):6.1f}")
