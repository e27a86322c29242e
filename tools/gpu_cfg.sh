#!/bin/bash
# configs[2], [3], [4] at N=1 + measured DRAM traffic of the stand-alone kernels
mkdir -p gpurun_out
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/c2_n1.json 2> gpurun_out/c2_n1.err; tail -c 300 gpurun_out/c2_n1.err
timeout 600 python bench.py --config 3 --steps 10 > gpurun_out/c3_n1.json 2> gpurun_out/c3_n1.err; tail -c 300 gpurun_out/c3_n1.err
timeout 900 python tools/sweep_ncu.py gpurun_out/r02_sweep_ncu.jsonl > gpurun_out/sweep_ncu.log 2>&1; tail -3 gpurun_out/sweep_ncu.log | cut -c1-300
cp gpurun_out/r02_sweep_ncu.jsonl profiles/r02_sweep_ncu.jsonl 2>/dev/null
timeout 900 python bench.py --config 4 --steps 100 > gpurun_out/c4_n1.json 2> gpurun_out/c4_n1.err; tail -c 300 gpurun_out/c4_n1.err
for f in c2_n1 c3_n1 c4_n1; do head -c 400 gpurun_out/$f.json; echo; done
