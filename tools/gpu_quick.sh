#!/bin/bash
# quick round-trip: CTC/step parity subset, bench (device only), phase timing
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -q -x -k "step or ctc" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,2), 'us/step  isolated', round(d['roofline']['kernel_ms_isolated']*1e3,2))
PY
PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so timeout 300 python tools/phase_timing.py > gpurun_out/${TAG}_phase.txt 2>&1
head -11 gpurun_out/${TAG}_phase.txt
