#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or city100" > $O/pytest_wf.txt 2>&1
tail -4 $O/pytest_wf.txt
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --workload $wl > $O/bench_${wl}_$name.json 2> $O/bench_${wl}_$name.err; }
for pf in 0 1; do
  run pf$pf city100 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_pf$pf.so HPMVS_WF=1 HPMVS_WF_SPLIT=0
  run pf$pf plane8 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_pf$pf.so HPMVS_WF=1 HPMVS_WF_SPLIT=0
  run pf$pf plane8x100k HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_pf$pf.so HPMVS_WF=1 HPMVS_WF_SPLIT=0
done
run auto city100 A=1
run auto plane8 A=1
HPMVS_WF_SPLIT=0 HPMVS_LIB=$PWD/hpmvs_b200/variants/libwf_pf1.so python scripts/wf_roundlog.py city100 $O/roundlog_city100_pf1.csv > $O/roundlog_city100_pf1.txt 2>&1
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))" 2>&1 | tail -1; done
awk 'NR<=3 || NR%4==0' $O/roundlog_city100_pf1.txt
