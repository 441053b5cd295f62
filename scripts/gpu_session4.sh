set -x
mkdir -p gpurun_out/s4
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config1 or overlapping" > gpurun_out/s4/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/s4/tests.txt
tail -4 gpurun_out/s4/tests.txt
for f in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --inflight $f > gpurun_out/s4/bench_plane8_f$f.json 2> gpurun_out/s4/bench_plane8_f$f.err
  timeout 300 python bench.py --steps 6 --warmup 3 --inflight $f --workload plane8x100k --cpu-sample 2000 > gpurun_out/s4/bench_100k_f$f.json 2> gpurun_out/s4/bench_100k_f$f.err
done
python - <<PY
import json
for n in ("bench_plane8_f1", "bench_plane8_f2", "bench_100k_f1", "bench_100k_f2"):
    try:
        d = json.load(open("gpurun_out/s4/%s.json" % n))
        print(n, "value %.0f ms %.2f e2e %.0f cpu %.0f (%s) frac %.4f ncc %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["roofline"]["frac"], d["roofline_ncc"]["frac"]))
    except Exception as e:
        print(n, "FAILED", e)
PY
