#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
r() { name=$1; shift; timeout 400 python bench.py --no-cpu --no-ncc "$@" > $O/$name.json 2>$O/$name.err; python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print('$name: n %d value %.0f e2e %.0f ms %.2f'%(d['run']['patches_per_step_this_rank'], d['value'],d['e2e']['value'],d['ms_per_step']))"; }
r sw8_if8 --sim-world 8 --inflight 8 --steps 32
r sw8_if12 --sim-world 8 --inflight 12 --steps 36
r sw8_if16 --sim-world 8 --inflight 16 --steps 32
r sw4_if12 --sim-world 4 --inflight 12 --steps 24
r city100_if8 --inflight 8 --steps 16
r city100_if12 --inflight 12 --steps 24
r plane8_if8 --workload plane8 --inflight 8 --steps 32
r plane8_if16 --workload plane8 --inflight 16 --steps 32
HPMVS_WF=0 r plane8_if8_persistent --workload plane8 --inflight 8 --steps 32
r plane8x100k_if8 --workload plane8x100k --inflight 8 --steps 8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
