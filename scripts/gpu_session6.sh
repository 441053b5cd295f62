mkdir -p gpurun_out/s6
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s6/tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s6/tests.txt
for pk in 0 1; do for f in 1 2; do
  HPMVS_PARKED=$pk timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 64 --inflight $f > gpurun_out/s6/p8_pk${pk}_f$f.json 2> gpurun_out/s6/p8_pk${pk}_f$f.err
  python -c "
import json; d=json.load(open('gpurun_out/s6/p8_pk${pk}_f$f.json')); print('parked=$pk inflight=$f', 'ms %.2f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
done; done
