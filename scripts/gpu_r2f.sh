#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront" > $O/pytest_wf.txt 2>&1
tail -4 $O/pytest_wf.txt
run() { # name workload env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --workload $wl > $O/bench_${wl}_$name.json 2> $O/bench_${wl}_$name.err
}
for p in 1 2 4 8; do
  run p$p city100 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=$p
  run p$p plane8 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=$p
done
run p4c32 city100 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4 CUDA_DEVICE_MAX_CONNECTIONS=32
run p8c32 city100 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=8 CUDA_DEVICE_MAX_CONNECTIONS=32
run p8s city100 HPMVS_WF=1 HPMVS_WF_SPLIT=1 HPMVS_WF_PARTS=8
run p4 plane8x100k HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4
run p8 plane8x100k HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=8
HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=8 timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --inflight 1 > $O/bench_city100_p8_if1.json 2> $O/bench_city100_p8_if1.err
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))" 2>&1 | tail -1; done
