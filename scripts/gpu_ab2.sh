# A/B on one box: libhp_base.so (committed state) vs the working tree; parity tests on the working tree first
mkdir -p gpurun_out/ab2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_next_rows.py -m gpu -q -x > gpurun_out/ab2/tests.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/ab2/tests.txt
for v in base new base new; do
  if [ $v = base ]; then export HPMVS_LIB=$PWD/hpmvs_b200/libhp_base.so; else unset HPMVS_LIB; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 64 --inflight 1 > gpurun_out/ab2/p8f1_$v.json 2>/dev/null
  timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 64 --inflight 2 > gpurun_out/ab2/p8f2_$v.json 2>/dev/null
  timeout 200 python bench.py --steps 5 --warmup 3 --cpu-sample 64 --inflight 2 --workload plane8x100k > gpurun_out/ab2/k100_$v.json 2>/dev/null
  python - <<PY
import json
def g(f, k="ms_per_step"):
    try:
        d = json.load(open(f)); return "%.2f" % d[k] if k else "%.4f" % d["roofline_ncc"]["frac"]
    except Exception as e: return "ERR"
print("$v", "plane8 f1", g("gpurun_out/ab2/p8f1_$v.json"), "f2", g("gpurun_out/ab2/p8f2_$v.json"), "100k f2", g("gpurun_out/ab2/k100_$v.json"), "ncc frac", g("gpurun_out/ab2/p8f1_$v.json", None))
PY
done
