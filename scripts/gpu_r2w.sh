#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
r() { name=$1; shift; timeout 400 python bench.py --no-cpu "$@" > $O/$name.json 2>$O/$name.err; python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); r=d.get('roofline_ncc') or {}; print('$name: value %.0f e2e %.0f ms %.2f launches %d ncc ms %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'], d['gpu_launches'], r.get('launch_ms',0)))"; }
for v in 0 1 0 1; do
HPMVS_LIB=$PWD/hpmvs_b200/variants/libvar$v.so r city100_var$v --steps 20 --warmup 5
done
for v in 0 1; do
HPMVS_LIB=$PWD/hpmvs_b200/variants/libvar$v.so r plane8_var$v --workload plane8 --steps 20 --warmup 5
HPMVS_LIB=$PWD/hpmvs_b200/variants/libvar$v.so HPMVS_WF=0 r plane8_persistent_var$v --workload plane8 --steps 20 --warmup 5 --inflight 2
done
