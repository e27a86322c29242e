#!/bin/bash
# phase timing of several -DPGASR_TIMING build variants in one gpurun call: tools/gpu_variants.sh TAG lib1 lib2 ...
mkdir -p gpurun_out
TAG=$1; shift
for L in "$@"; do
  N=$(basename $L .so | sed 's/libpgasr_b200_timing//; s/[^A-Za-z0-9_]/-/g')
  PGASR_LIB=$L timeout 150 python tools/phase_timing.py > gpurun_out/${TAG}${N}.txt 2>&1
  echo "== $N rc=$?"; sed -n '2,5p' gpurun_out/${TAG}${N}.txt | cut -c1-150; grep -A1 "pg sampling" gpurun_out/${TAG}${N}.txt | head -2 | cut -c1-200; tail -4 gpurun_out/${TAG}${N}.txt | cut -c1-200
done
