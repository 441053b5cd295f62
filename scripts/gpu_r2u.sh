#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_city100_reference.json 2> $O/bench_city100_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench_city100.json 2> $O/bench_city100.err
python bench.py --workload plane8 --steps 20 --warmup 5 > $O/bench_plane8.json 2> $O/bench_plane8.err
python bench.py --workload plane8 --impl reference --steps 20 --warmup 5 > $O/bench_plane8_reference.json 2> $O/bench_plane8_reference.err
python bench.py --workload plane8x100k --no-cpu --steps 12 --warmup 5 > $O/bench_plane8x100k.json 2> $O/bench_plane8x100k.err
HPMVS_NCC_TMA=1 python bench.py --no-cpu --steps 4 > $O/bench_city100_ncc_staged.json 2> $O/bench_city100_ncc_staged.err
HPMVS_NCC_TMA=1 python bench.py --workload plane8 --no-cpu --steps 4 > $O/bench_plane8_ncc_staged.json 2> $O/bench_plane8_ncc_staged.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline_ncc') or {}; print(' value %.0f e2e %.0f ms %.2f roof %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],(d.get('roofline') or {}).get('frac',0)), 'ncc frac', r.get('frac'), 'ncc ms', r.get('launch_ms'), 'staged', r.get('tma_staged_windows'), (d.get('cpu_baseline') or {}).get('value'))" 2>&1 | tail -1; done
ncu --set full --clock-control none --import-source on -k regex:ncc_kernel -s 1 -c 1 -o $O/ncc_plain python scripts/ncu_step.py city100 ncc > $O/ncu_ncc_plain.log 2>&1
HPMVS_NCC_TMA=1 ncu --set full --clock-control none --import-source on -k regex:ncc_kernel -s 1 -c 1 -o $O/ncc_staged python scripts/ncu_step.py city100 ncc > $O/ncu_ncc_staged.log 2>&1
HPMVS_WF=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --print-units base -k regex:wf_ --csv --log-file $O/step_kernels_city100.csv python scripts/ncu_step.py city100 > $O/ncu_step.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2u/step_kernels_city100.csv") if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0]; agg[k][r[mi]] += float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum": cnt[k] += 1
tot = collections.defaultdict(float)
for k, m in agg.items():
    print(f"{k:44s} launches {cnt[k]:5d}", {a: round(b / 1e6, 3) for a, b in m.items()})
    for a, b in m.items(): tot[a] += b
print("TOTAL", {a: b for a, b in tot.items()})
PY
