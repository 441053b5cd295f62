#!/bin/bash
# round 2, GPU session A: tests on the round-2 tree, city100 bench (both arms), cycle accounting on city100
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_city100_reference.json 2> $O/bench_city100_reference.err
python bench.py > $O/bench_city100.json 2> $O/bench_city100.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
tail -3 $O/pytest_gpu.txt
HPMVS_LIB=$PWD/hpmvs_b200/libhpmvs_b200_prof.so HPMVS_PROFILE_PRINT=1 python bench.py --steps 3 --no-cpu --no-ncc --inflight 1 > $O/prof_city100.json 2> $O/prof_city100.err
python bench.py --workload plane8 > $O/bench_plane8.json 2> $O/bench_plane8.err
cat $O/bench_city100.json | head -c 1500
