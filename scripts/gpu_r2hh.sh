#!/bin/bash
O=gpurun_out/r2hh; mkdir -p $O
r() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-ncc --steps 20 --warmup 5 "$@" > $O/$name.json 2>$O/$name.err; python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print('$name: value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"; }
HPMVS_WF_SAMPLER_CTAS=7 r city100_ctas7
r city100_occ
HPMVS_WF_SAMPLER_CTAS=5 r city100_ctas5
HPMVS_WF_SAMPLER_CTAS=7 r plane8_ctas7 --workload plane8
r plane8_occ --workload plane8
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or city100" 2>&1 | tail -2
