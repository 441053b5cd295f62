#!/bin/bash
# round 2, GPU session B: first run of the wavefront kernels (correctness, then timing against the persistent kernel)
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront" > $O/pytest_wf.txt 2>&1
tail -15 $O/pytest_wf.txt
for cfg in "0 1" "1 1" "1 0"; do set -- $cfg
  HPMVS_WF=$1 HPMVS_WF_SPLIT=$2 timeout 600 python bench.py --no-cpu --no-ncc --steps 6 > $O/bench_city100_wf$1_split$2.json 2> $O/bench_city100_wf$1_split$2.err
  HPMVS_WF=$1 HPMVS_WF_SPLIT=$2 timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --workload plane8 > $O/bench_plane8_wf$1_split$2.json 2> $O/bench_plane8_wf$1_split$2.err
done
HPMVS_WF=1 timeout 600 python bench.py --no-cpu --no-ncc --steps 4 --workload plane8x100k > $O/bench_plane8x100k_wf1.json 2> $O/bench_plane8x100k_wf1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))" 2>&1 | tail -1; done
