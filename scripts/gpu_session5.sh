# final-state profiles of round 1: launch list of the default bench command + one full capture of the fused kernel
set -x
mkdir -p gpurun_out/s5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5/launches.csv python bench.py --steps 4 --warmup 3 --cpu-sample 256 > gpurun_out/s5/launches_bench.log 2>&1
tail -2 gpurun_out/s5/launches_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -s 3 -c 1 -o gpurun_out/s5/opt_full python bench.py --steps 2 --warmup 3 --cpu-sample 256 --inflight 1 > gpurun_out/s5/ncu_opt.log 2>&1
tail -3 gpurun_out/s5/ncu_opt.log | cut -c1-200
ls -la gpurun_out/s5
