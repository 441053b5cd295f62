#!/bin/bash
# round 2: validation of the final tree on one GPU
O=gpurun_out/r2ff; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-ncc > $O/bench_city100.json 2> $O/bench_city100.err
python bench.py --workload plane8 --steps 20 --warmup 5 --no-ncc --no-cpu > $O/bench_plane8.json 2> $O/bench_plane8.err
python bench.py --workload plane8 --steps 8 --warmup 3 --no-ncc --no-cpu --inflight 1 > $O/bench_plane8_inflight1.json 2> $O/bench_plane8_inflight1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f launches %d warm %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches'],d['run']['warmup_launches']), (d.get('cpu_baseline') or {}).get('value'))" 2>&1 | tail -1; done
