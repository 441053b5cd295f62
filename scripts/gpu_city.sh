mkdir -p gpurun_out/city
timeout 900 python bench.py --workload city100 --steps 4 --warmup 3 > gpurun_out/city/bench_city100.json 2> gpurun_out/city/bench_city100.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/city/bench_city100.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'], 'frac', d['roofline']['frac'], 'n', d['config']['patches_per_step_per_gpu'], d['config']['optimized_per_step'])"
timeout 600 python bench.py --workload city100 --impl reference --steps 3 --warmup 1 > gpurun_out/city/ref_city100.json 2> gpurun_out/city/ref_city100.err
echo "rc=$?"; head -c 400 gpurun_out/city/ref_city100.json
