#!/bin/bash
# step time of the two roles alone and together at the headline shape, one box (diagnosis)
mkdir -p gpurun_out
TAG=${1:-roles}
for W in "1 1" "0 1" "1 0"; do
  set -- $W
  for E in "" "PGASR_NO_OVERLAP=1"; do
    env $E timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 400 --warmup 50 --w-pg $1 --w-ctc $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('w_pg $1 w_ctc $2 $E:', round(d['ms_per_step']*1e3,2), 'us/step  isolated', round(d['roofline']['kernel_ms_isolated']*1e3,2))" | tee -a gpurun_out/${TAG}.txt
  done
done
