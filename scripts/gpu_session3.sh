# GPU session 3: tests, bench with the stand-alone scoring kernel figure, ncu captures of both kernels
set -x
mkdir -p gpurun_out/s3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/s3/tests.txt
tail -3 gpurun_out/s3/tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s3/bench_plane8.json 2> gpurun_out/s3/bench_plane8.err
python -c "
import json; d=json.load(open('gpurun_out/s3/bench_plane8.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac']); print('ncc', d['roofline_ncc'])"
# ncu: the scoring kernel, one launch, full set
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncc_kernel -s 2 -c 1 -o gpurun_out/s3/ncc_full python bench.py --steps 2 --warmup 3 --cpu-sample 256 > gpurun_out/s3/ncu_ncc.log 2>&1
tail -3 gpurun_out/s3/ncu_ncc.log
