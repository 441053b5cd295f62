#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
for sp in 0 1; do
HPMVS_WF_SPLIT=$sp python scripts/wf_roundlog.py city100 $O/roundlog_city100_split$sp.csv > $O/roundlog_city100_split$sp.txt 2>&1
HPMVS_WF_SPLIT=$sp python scripts/wf_roundlog.py plane8 $O/roundlog_plane8_split$sp.csv > $O/roundlog_plane8_split$sp.txt 2>&1
done
cat $O/roundlog_city100_split0.txt
HPMVS_WF=2 HPMVS_WF_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wf_ -c 700 --csv --log-file $O/launches_city100_wf2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-ncc --inflight 1 > $O/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2c/launches_city100_wf2.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
seq = [(r[ki].split("(")[0][-40:], float(r[vi].replace(",", ""))) for r in rows[1:]]
agg = collections.OrderedDict()
for i, (k, v) in enumerate(seq):
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    print(f"{k:42s} n={len(v):4d} mean {sum(v)/len(v)/1e3:9.1f} us  first10 {[round(x/1e3) for x in v[:10]]}")
PY
