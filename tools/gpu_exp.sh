#!/bin/bash
# phase timing of experimental -D variants of the timing build: tools/gpu_exp.sh tag1 tag2 ...
mkdir -p gpurun_out
for t in "$@"; do
  echo "=== $t"
  PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing_$t.so timeout 300 python tools/phase_timing.py > gpurun_out/exp_$t.txt 2>&1
  sed -n '2,11p' gpurun_out/exp_$t.txt
done
