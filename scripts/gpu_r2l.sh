#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --no-ncc > $O/bench_city100_2gpu.json 2> $O/bench_city100_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --no-cpu --no-ncc --inflight 1 > $O/bench_city100_2gpu_if1.json 2> $O/bench_city100_2gpu_if1.err
for f in $O/bench_city100_2gpu*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'])" 2>&1 | tail -1; done
tail -n 3 $O/bench_city100_2gpu.err
# configs[4] family, scaled probe on GPU 0: 40 views 4K, 40k seed points (render + upload rate, HBM per view)
HPMVS_CITY500_VIEWS=40 HPMVS_CITY500_SEEDS=40000 timeout 900 python bench.py --workload city500_4k --steps 3 --no-ncc > $O/bench_city4k_40views.json 2> $O/bench_city4k_40views.err
python -c "
import json
d=json.loads(open('$O/bench_city4k_40views.json').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['run']['scene_upload_s'], d['run']['hbm_used_gb'], d['run']['status_histogram_rank0'], d['config']['patches_per_step'])"
tail -n 3 $O/bench_city4k_40views.err
