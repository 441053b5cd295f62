set -x
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/s1/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/s1/tests.txt
tail -5 gpurun_out/s1/tests.txt
for cfg in default 3,9,23 3,10,23 3,10,26; do
  if [ $cfg = default ]; then unset HPMVS_CONFIG; else export HPMVS_CONFIG=$cfg; fi
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s1/bench_plane8_$cfg.json 2> gpurun_out/s1/bench_plane8_$cfg.err
  python -c "import json,sys; d=json.load(open('gpurun_out/s1/bench_plane8_$cfg.json')); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])"
done
for cfg in default 3,10,26; do
  if [ $cfg = default ]; then unset HPMVS_CONFIG; unset HPMVS_PARKED; else export HPMVS_CONFIG=$cfg; export HPMVS_PARKED=0; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --workload plane8x100k > gpurun_out/s1/bench_100k_$cfg.json 2> gpurun_out/s1/bench_100k_$cfg.err
  python -c "import json,sys; d=json.load(open('gpurun_out/s1/bench_100k_$cfg.json')); print('100k $cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])"
done
unset HPMVS_CONFIG; export HPMVS_PARKED=0
timeout 300 python bench.py --steps 5 --warmup 3 --workload plane8x100k > gpurun_out/s1/bench_100k_resident.json 2>/dev/null
python -c "import json,sys; d=json.load(open('gpurun_out/s1/bench_100k_resident.json')); print('100k resident 2,10,32', d['value'], d['ms_per_step'])"
