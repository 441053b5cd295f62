#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -3
for sw in 8 4 2; do for wf in 0 1; do for f in 4 8; do
HPMVS_WF=$wf timeout 300 python bench.py --no-cpu --no-ncc --sim-world $sw --inflight $f --steps 16 > $O/b_sw${sw}_wf${wf}_if$f.json 2>/dev/null
python -c "
import json
d=json.loads(open('$O/b_sw${sw}_wf${wf}_if$f.json').read().strip().splitlines()[-1]); print('sim-world $sw wf $wf inflight $f: n %d value %.0f e2e %.0f ms %.2f'%(d['run']['patches_per_step_this_rank'], d['value'],d['e2e']['value'],d['ms_per_step']))"
done; done; done
