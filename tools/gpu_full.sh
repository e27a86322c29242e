#!/bin/bash
# full GPU suite + headline bench (+ A/B switches given as env assignments in $2..)
TAG=${1:-f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_s20.json 2> gpurun_out/${TAG}_bench_s20.err
PGASR_NO_OVERLAP=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_noov.json 2> gpurun_out/${TAG}_bench_noov.err
for f in bench bench_s20 bench_noov; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,2), 'us/step  isolated', round(d['roofline']['kernel_ms_isolated']*1e3,2), 'frac', round(d['roofline']['frac'],4), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print('$f failed', e)
PY
done
