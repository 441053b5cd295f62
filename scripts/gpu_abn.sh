# A/B/n on ONE box: for every variant library hpmvs_b200/libhp_<name>.so run the bit-exact GPU parity tests once and the
# bench on plane8 (10k) and plane8x100k; usage: bash scripts/gpu_abn.sh <tag> name1 name2 ...
tag=$1; shift
mkdir -p gpurun_out/$tag
for v in "$@"; do
  export HPMVS_LIB=$PWD/hpmvs_b200/libhp_$v.so
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_exact or reference_fixture or golden" > gpurun_out/$tag/tests_$v.txt 2>&1
  t=$(tail -1 gpurun_out/$tag/tests_$v.txt)
  for rep in 1 2; do
    timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 64 > gpurun_out/$tag/p8_${v}_$rep.json 2> gpurun_out/$tag/p8_${v}_$rep.err
  done
  timeout 300 python bench.py --steps 4 --warmup 3 --cpu-sample 64 --workload plane8x100k > gpurun_out/$tag/k100_$v.json 2> gpurun_out/$tag/k100_$v.err
  python - <<PY
import json
def ms(f):
    try: return "%.2f" % json.load(open(f))["ms_per_step"]
    except Exception as e: return "ERR"
print("$v", "| tests:", "$t", "| plane8 ms", ms("gpurun_out/$tag/p8_${v}_1.json"), ms("gpurun_out/$tag/p8_${v}_2.json"), "| 100k ms", ms("gpurun_out/$tag/k100_$v.json"))
PY
done
