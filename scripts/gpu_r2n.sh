#!/bin/bash
# round 2, 8-GPU session: strong scaling of the city100 batch (octree sub-trees dealt to the ranks) and configs[4] once
O=gpurun_out/r2n; mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 > $O/gpus.txt; nproc >> $O/gpus.txt
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-ncc > $O/bench_city100_${n}gpu.json 2> $O/bench_city100_${n}gpu.err
done
for f in 4 6 8; do timeout 600 python bench.py --no-cpu --no-ncc --inflight $f --steps 16 > $O/bench_city100_1gpu_if$f.json 2>/dev/null; done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload city500_4k --steps 3 --no-ncc --no-cpu > $O/bench_city500_4k_8gpu.json 2> $O/bench_city500_4k_8gpu.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'], d['run']['scene_upload_s'], d['run']['hbm_used_gb'], (d.get('cpu_baseline') or {}).get('value'), d['run']['status_histogram_rank0'])" 2>&1 | tail -1; done
tail -n 4 $O/bench_city500_4k_8gpu.err
