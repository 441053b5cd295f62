#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-ncc --no-cpu > $O/bench_city100_${n}gpu.json 2> $O/bench_city100_${n}gpu.err
done
for f in $O/bench_*gpu.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'], d['run']['device_ms_per_step_by_rank'], d['run']['shards']['subtrees'], d['run']['shards']['seeds_per_rank'])" 2>&1 | tail -1; done
tail -n 3 $O/bench_city100_8gpu.err
