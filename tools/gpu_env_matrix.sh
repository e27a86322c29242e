#!/bin/bash
# bench (device only) under combinations of the A/B environment switches, one box
mkdir -p gpurun_out
TAG=${1:-envm}
for r in 1 2; do
for E in "PGASR_LANES=3" "PGASR_LANES=3 PGASR_NO_PDL=1" "PGASR_LANES=2" "PGASR_LANES=2 PGASR_NO_PDL=1" "PGASR_LANES=4" "PGASR_LANES=4 PGASR_NO_PDL=1"; do
  for S in 400 20; do
  env $E timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps $S --warmup 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$E steps $S:', round(d['ms_per_step']*1e3,2), 'us/step')" | tee -a gpurun_out/${TAG}.txt
  done
done
done
