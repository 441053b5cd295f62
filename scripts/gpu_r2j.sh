#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
tail -6 $O/pytest_gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/pipeline_multigpu.py 1000 > $O/pipeline_2gpu.json 2> $O/pipeline_2gpu.err
tail -c 1500 $O/pipeline_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/bench_city100_2gpu.json 2> $O/bench_city100_2gpu.err
python bench.py > $O/bench_city100_1gpu.json 2> $O/bench_city100_1gpu.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['run']['patches_gathered_kept'], d['run']['e2e_gather_dedup_ms_per_step'], d['run'].get('shards'))" 2>&1 | tail -1; done
tail -3 $O/*.err | tail -20
