#!/bin/bash
# round 2, final 8-GPU session: strong scaling of configs[3] at N = 8 / 4 / 2 and configs[4] (500 views 4K) once
O=gpurun_out/r2q; mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 > $O/gpus.txt; nproc >> $O/gpus.txt; free -g | head -2 >> $O/gpus.txt
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-ncc > $O/bench_city100_${n}gpu.json 2> $O/bench_city100_${n}gpu.err
done
timeout 600 python bench.py --no-ncc > $O/bench_city100_1gpu.json 2> $O/bench_city100_1gpu.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload city500_4k --steps 4 --no-ncc --no-cpu > $O/bench_city500_4k_8gpu.json 2> $O/bench_city500_4k_8gpu.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'], d['run']['scene_upload_s'], d['run']['hbm_used_gb'], (d.get('cpu_baseline') or {}).get('value'), d['run']['status_histogram_rank0'], d['run']['device_ms_per_step_by_rank'])" 2>&1 | tail -1; done
tail -n 4 $O/bench_city500_4k_8gpu.err
