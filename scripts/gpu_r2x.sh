#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
r() { name=$1; shift; timeout 400 python bench.py --no-cpu "$@" > $O/$name.json 2>$O/$name.err; python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); r=d.get('roofline_ncc') or {}; print('$name: value %.0f e2e %.0f ms %.2f ncc ms %.3f frac %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'], r.get('launch_ms',0), r.get('frac',0)))"; }
for v in base new base new; do
if [ $v = base ]; then export HPMVS_LIB=$PWD/hpmvs_b200/variants/libbase.so; else unset HPMVS_LIB; fi
r city100_$v --steps 20 --warmup 5
done
for v in base new; do
if [ $v = base ]; then export HPMVS_LIB=$PWD/hpmvs_b200/variants/libbase.so; else unset HPMVS_LIB; fi
r plane8_$v --workload plane8 --steps 20 --warmup 5
done
