#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
for m in 2 1; do for t in 2 4; do
HPMVS_WF=$m python scripts/wf_concurrency_probe.py city100 $t 2>&1 | tail -1 | tee -a $O/probe.txt
done; done
HPMVS_WF=2 CUDA_DEVICE_MAX_CONNECTIONS=32 python scripts/wf_concurrency_probe.py city100 4 2>&1 | tail -1 | tee -a $O/probe.txt
HPMVS_WF=0 python scripts/wf_concurrency_probe.py city100 2 2>&1 | tail -1 | tee -a $O/probe.txt
