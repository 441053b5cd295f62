# bulk-async (TMA engine) staging of the parked kernel: parity first (short timeout: a wrong transaction count would hang), then A/B
mkdir -p gpurun_out/tma
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "parked or overlapping" > gpurun_out/tma/tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tma/tests.txt
for v in notma tma notma tma; do
  if [ $v = notma ]; then export HPMVS_LIB=$PWD/hpmvs_b200/libhp_notma.so; else unset HPMVS_LIB; fi
  timeout 200 python bench.py --steps 6 --warmup 3 --workload plane8x100k --cpu-sample 64 --inflight 1 > gpurun_out/tma/k100_$v.json 2> gpurun_out/tma/k100_$v.err
  timeout 200 python bench.py --steps 4 --warmup 3 --workload city100 --cpu-sample 64 --inflight 1 > gpurun_out/tma/city_$v.json 2> gpurun_out/tma/city_$v.err
  python - <<PY
import json
def ms(f):
    try: return "%.2f" % json.load(open(f))["ms_per_step"]
    except Exception as e: return "ERR"
print("$v", "100k ms", ms("gpurun_out/tma/k100_$v.json"), "city100 ms", ms("gpurun_out/tma/city_$v.json"))
PY
done
