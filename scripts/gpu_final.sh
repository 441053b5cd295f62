# what the driver runs at round end, plus the secondary workloads: gpu tests, smoke(), default bench, reference arm
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/tests.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final/tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final/smoke.txt
timeout 300 python bench.py --impl reference --gpus 1 > gpurun_out/final/ref_plane8.json 2> gpurun_out/final/ref_plane8.err; echo "ref rc=$?"
timeout 300 python bench.py --gpus 1 > gpurun_out/final/bench_plane8.json 2> gpurun_out/final/bench_plane8.err; echo "bench rc=$?"
timeout 300 python bench.py --workload plane8x100k --steps 6 > gpurun_out/final/bench_100k.json 2> gpurun_out/final/bench_100k.err
python - <<PY
import json
r = json.load(open("gpurun_out/final/ref_plane8.json"))
for n in ("bench_plane8", "bench_100k"):
    d = json.load(open("gpurun_out/final/%s.json" % n))
    print(n, "value %.0f ms %.2f e2e %.0f cpu_baseline %.0f (%s, %d cores) frac %.4f ncc %.4f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], d["roofline"]["frac"], d["roofline_ncc"]["frac"], d["gpu_launches"]), d["clocks"])
print("reference arm %.0f patches/s on %d cores" % (r["value"], r["cpu_baseline"]["cores"]))
PY
