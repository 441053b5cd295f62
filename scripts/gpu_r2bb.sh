#!/bin/bash
O=gpurun_out/r2bb; mkdir -p $O
timeout 500 python bench.py --workload fountain11 --steps 2 > $O/bench_fountain11.json 2> $O/bench_fountain11.err
tail -c 1800 $O/bench_fountain11.json; tail -n 3 $O/bench_fountain11.err
timeout 900 python bench.py --workload fountain11 --impl reference --steps 1 > $O/bench_fountain11_reference.json 2> $O/bench_fountain11_reference.err
tail -c 900 $O/bench_fountain11_reference.json; tail -n 3 $O/bench_fountain11_reference.err
