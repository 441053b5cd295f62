#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront" > $O/pytest_wf.txt 2>&1
tail -4 $O/pytest_wf.txt
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-ncc --steps 6 > $O/bench_city100_$name.json 2> $O/bench_city100_$name.err
  env "$@" timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --workload plane8 > $O/bench_plane8_$name.json 2> $O/bench_plane8_$name.err
}
run wf1_p1 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=1
run wf1_p2 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=2
run wf1_p4 HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4
run wf1_p4s HPMVS_WF=1 HPMVS_WF_SPLIT=1 HPMVS_WF_PARTS=4
HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4 timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --inflight 1 > $O/bench_city100_wf1_p4_if1.json 2> $O/bench_city100_wf1_p4_if1.err
HPMVS_WF=1 HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=4 timeout 600 python bench.py --no-cpu --no-ncc --steps 6 --inflight 3 > $O/bench_city100_wf1_p4_if3.json 2> $O/bench_city100_wf1_p4_if3.err
HPMVS_WF=1 HPMVS_WF_SPLIT=0 timeout 600 python bench.py --no-cpu --no-ncc --steps 4 --workload plane8x100k > $O/bench_plane8x100k_wf1_p4.json 2> $O/bench_plane8x100k_wf1_p4.err
HPMVS_WF_SPLIT=0 HPMVS_WF_PARTS=1 python scripts/wf_roundlog.py city100 $O/roundlog_city100_p1.csv > $O/roundlog_city100_p1.txt 2>&1
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))" 2>&1 | tail -1; done
