#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --workload $wl > $O/bench_${wl}_$name.json 2> $O/bench_${wl}_$name.err; }
for m in 4 5; do for p in 1 2 4; do
  run cta${m}_p$p city100 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_cta$m.so HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=$p
done; done
run cta5_p4 plane8 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_cta5.so HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4
run cta5_p1 plane8 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_cta5.so HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=1
HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_cta5.so HPMVS_WF=1 python scripts/wf_concurrency_probe.py city100 2 2>&1 | tail -1 | tee -a $O/probe.txt
HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_cta5.so HPMVS_WF=1 python scripts/wf_concurrency_probe.py city100 4 2>&1 | tail -1 | tee -a $O/probe.txt
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))" 2>&1 | tail -1; done
