#!/bin/bash
# round 2, call B: block workers parity + A/B
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
PGASR_NO_BW=1 timeout 900 python -m pytest tests -m gpu -q -k "step or ctc" > gpurun_out/r2b_pytest_nobw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_nobw.log
tail -3 gpurun_out/r2b_pytest_nobw.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_bw.json 2> gpurun_out/r2b_bench_bw.err
PGASR_NO_BW=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_nobw.json 2> gpurun_out/r2b_bench_nobw.err
for sw in "20 5" "20 50" "44 5" "100 5"; do set -- $sw; timeout 300 python bench.py --steps $1 --warmup $2 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_s$1_w$2.json 2> gpurun_out/r2b_bench_s$1_w$2.err; done
timeout 300 python bench.py --batch 74 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_b74.json 2>/dev/null
timeout 300 python bench.py --config 2 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_c2.json 2>/dev/null
PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so timeout 300 python tools/phase_timing.py > gpurun_out/r2b_phase_bw.txt 2>&1
PGASR_NO_BW=1 PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so timeout 300 python tools/phase_timing.py > gpurun_out/r2b_phase_nobw.txt 2>&1
timeout 900 python tools/fuzz_step.py 350 11 > gpurun_out/r2b_fuzz350.log 2>&1; tail -2 gpurun_out/r2b_fuzz350.log
for f in bw nobw s20_w5 s20_w50 s44_w5 s100_w5 b74 c2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms_isolated'], (d.get('e2e') or {}).get('value'))
"; done
head -12 gpurun_out/r2b_phase_bw.txt
