#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k tma > $O/sanitizer_tma.txt 2>&1
grep -v "^$" $O/sanitizer_tma.txt | grep -E "=====|Error|error|at |by " | head -40
r() { name=$1; shift; timeout 400 python bench.py --no-cpu --no-ncc "$@" > $O/$name.json 2>$O/$name.err; python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print('$name: n %d value %.0f e2e %.0f ms %.2f'%(d['run']['patches_per_step_this_rank'], d['value'],d['e2e']['value'],d['ms_per_step']))"; }
for f in 4 6 8 12; do r city100_if$f --inflight $f --steps 20 --warmup 5; done
for f in 4 8 12; do r sw2_if$f --sim-world 2 --inflight $f --steps 20 --warmup 5; done
for f in 4 8 12; do r sw4_if$f --sim-world 4 --inflight $f --steps 20 --warmup 5; done
for f in 4 8 12 16; do r sw8_if$f --sim-world 8 --inflight $f --steps 20 --warmup 5; done
for f in 4 8 16; do r plane8_if$f --workload plane8 --inflight $f --steps 20 --warmup 5; done
