# quick A/B: parity tests + bench lines for plane8 (10k) and plane8x100k; usage: bash scripts/gpu_ab.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out/$tag
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/$tag/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/$tag/tests.txt
tail -3 gpurun_out/$tag/tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 2000 > gpurun_out/$tag/bench_plane8.json 2> gpurun_out/$tag/bench_plane8.err
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 2000 --workload plane8x100k > gpurun_out/$tag/bench_100k.json 2> gpurun_out/$tag/bench_100k.err
HPMVS_PARKED=0 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 2000 --workload plane8x100k > gpurun_out/$tag/bench_100k_resident.json 2> gpurun_out/$tag/bench_100k_resident.err
python - <<PY
import json
for n in ("bench_plane8", "bench_100k", "bench_100k_resident"):
    try:
        d = json.load(open("gpurun_out/$tag/%s.json" % n))
        print(n, "value %.0f ms %.2f e2e %.0f cpu %.0f (%s) frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["roofline"]["frac"]))
    except Exception as e:
        print(n, "FAILED", e)
PY
