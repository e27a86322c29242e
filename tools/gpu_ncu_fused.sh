#!/bin/bash
# one --set full capture (with source) of the fused kernel at the headline shape; TAG = output prefix
TAG=${1:-ncu}
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 -k regex:pg_ctc_fused -s 2 -c 1 -f -o gpurun_out/${TAG} \
    python tools/prof_step.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
