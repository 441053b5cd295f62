#!/bin/bash
# round 2, final multi-GPU record: strong scaling of configs[3] with the final settings, the driver's K / W
O=gpurun_out/r2v; mkdir -p $O
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-ncc > $O/bench_city100_${n}gpu.json 2> $O/bench_city100_${n}gpu.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-ncc > $O/bench_city100_1gpu.json 2> $O/bench_city100_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > $O/bench_city100_8gpu_reference.json 2> $O/bench_city100_8gpu_reference.err
for f in $O/bench_*gpu.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'], (d.get('cpu_baseline') or {}).get('value'), d['run']['device_ms_per_step_by_rank'])" 2>&1 | tail -1; done
tail -c 400 $O/bench_city100_8gpu_reference.json
