#!/bin/bash
# A/B of library builds on ONE box: tools/gpu_ab.sh TAG ROUNDS lib1 lib2 ...   (bench.py device-only, alternating)
mkdir -p gpurun_out
TAG=$1; R=$2; shift 2
for r in $(seq 1 $R); do
  for L in "$@"; do
    N=$(basename $L .so | sed 's/libpgasr_b200//; s/[^A-Za-z0-9_]/-/g')
    for S in 400 20; do
      PGASR_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps $S --warmup 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$TAG round $r lib[$N] steps $S:', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,2), 'us/step')" | tee -a gpurun_out/${TAG}_ab.txt
    done
  done
done
