#!/bin/bash
# final captures of the round: bench N=1 (400 steps and 20 steps), reference arm, phase timing, sampler ncu
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 200 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_s20_n1.json 2> gpurun_out/r02_bench_s20_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
for b in 74 128 148; do timeout 300 python bench.py --batch $b --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_b$b.json 2>/dev/null; done
timeout 300 python bench.py --regime peaky --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_peaky.json 2>/dev/null
PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so timeout 300 python tools/phase_timing.py > gpurun_out/r02_phase_timing.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:softmax_sample -s 1 -c 1 -f -o gpurun_out/r02_sampler python tools/prof_standalone.py > gpurun_out/p_ncu_sampler.log 2>&1
for f in r02_bench_n1 r02_bench_s20_n1 r02_bench_reference_arm r02_bench_b74 r02_bench_b128 r02_bench_b148 r02_bench_peaky; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,2), 'us/step', 'frac', round((d.get('roofline') or {}).get('frac',0),4), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('$f failed', e)
PY
done
