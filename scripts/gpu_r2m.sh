#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --no-ncc > $O/bench_city100_2gpu.json 2> $O/bench_city100_2gpu.err
for f in $O/bench_city100_2gpu*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'])" 2>&1 | tail -1; done
for f in 1 2 3 4; do python bench.py --no-cpu --no-ncc --inflight $f --steps 8 > $O/bench_city100_if$f.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('$O/bench_city100_if$f.json').read().strip().splitlines()[-1]); print('inflight $f value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"; done
