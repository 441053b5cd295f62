#!/bin/bash
# launch list of the bench command itself (host-driven rounds so that every kernel of a step is an ordinary launch)
O=gpurun_out/r2ii; mkdir -p $O
HPMVS_WF=2 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --print-units base -k regex:"wf_|optimize_kernel" -c 3500 --csv --log-file $O/launches_bench_city100.csv python bench.py --steps 2 --warmup 1 --inflight 1 --no-cpu --no-ncc > $O/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2ii/launches_bench_city100.csv") if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0]; agg[k] += float(r[vi].replace(",", "")); cnt[k] += 1
tot = sum(agg.values())
for k in agg: print(f"{k:44s} launches {cnt[k]:5d} time {agg[k]/1e6:9.3f} ms  share {100*agg[k]/tot:5.1f} %")
PY
