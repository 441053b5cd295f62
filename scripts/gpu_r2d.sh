#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
# full-population main-iteration rounds: skip the fill + 20 rounds (x6 kernels), then capture 2 rounds of every wavefront kernel
HPMVS_WF=2 HPMVS_WF_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:wf_ -s 121 -c 12 -o $O/wf_city100_split1 python bench.py --steps 1 --warmup 1 --no-cpu --no-ncc --inflight 1 > $O/ncu1.log 2>&1
HPMVS_WF=2 HPMVS_WF_SPLIT=0 ncu --set full --clock-control none --import-source on -k regex:wf_advance -s 20 -c 2 -o $O/wf_city100_split0 python bench.py --steps 1 --warmup 1 --no-cpu --no-ncc --inflight 1 > $O/ncu0.log 2>&1
ls -la $O
