#!/bin/bash
# round 2: final record run on one GPU with the final binary
O=gpurun_out/r2y; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_city100_reference.json 2> $O/bench_city100_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench_city100.json 2> $O/bench_city100.err
python bench.py --workload plane8 --steps 20 --warmup 5 > $O/bench_plane8.json 2> $O/bench_plane8.err
python bench.py --workload plane8x100k --no-cpu --steps 12 --warmup 5 > $O/bench_plane8x100k.json 2> $O/bench_plane8x100k.err
python bench.py --no-cpu --no-ncc --inflight 1 --steps 8 > $O/bench_city100_inflight1.json 2> $O/bench_city100_inflight1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline_ncc') or {}; print(' value %.0f e2e %.0f ms %.2f roof %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],(d.get('roofline') or {}).get('frac',0)), 'ncc frac', r.get('frac'), 'ncc ms', r.get('launch_ms'), (d.get('cpu_baseline') or {}).get('value'), d.get('gpu_launches'))" 2>&1 | tail -1; done
ncu --set full --clock-control none -k regex:wf_eval -s 40 -c 1 -o $O/wf_eval_final python scripts/ncu_step.py city100 > $O/ncu_eval.log 2>&1
