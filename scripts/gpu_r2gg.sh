#!/bin/bash
O=gpurun_out/r2gg; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-ncc > $O/bench_city100_2gpu.json 2> $O/bench_city100_2gpu.err
python -c "
import json
d=json.loads(open('$O/bench_city100_2gpu.json').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['device_ms_per_step_by_rank'], d['gpu_launches'], (d.get('cpu_baseline') or {}).get('value'))"
tail -n 2 $O/bench_city100_2gpu.err
