#!/bin/bash
# final captures of the round (one GPU): full parity suite, bench lines, phase timing, launch list, --set full of the fused kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log; tail -3 gpurun_out/r02_pytest.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 200 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_s20_n1.json 2> gpurun_out/r02_bench_s20_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 300 python bench.py --reward ed_to_go --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_reward_to_go_n1.json 2>/dev/null
timeout 300 python bench.py --config 2 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_config2_n1.json 2>/dev/null
for b in 74 128 148; do timeout 300 python bench.py --batch $b --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_b$b.json 2>/dev/null; done
timeout 300 python bench.py --regime peaky --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_peaky.json 2>/dev/null
PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so timeout 300 python tools/phase_timing.py > gpurun_out/r02_phase_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/p_ncu_launches.log 2>&1
timeout 300 python tools/prof_step.py > gpurun_out/p_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 \
    -k regex:pg_ctc_fused -s 2 -c 1 -f -o gpurun_out/r02_fused python tools/prof_step.py > gpurun_out/p_ncu_fused.log 2>&1
for f in r02_bench_n1 r02_bench_s20_n1 r02_bench_reference_arm r02_bench_reward_to_go_n1 r02_bench_config2_n1 r02_bench_b74 r02_bench_b128 r02_bench_b148 r02_bench_peaky; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), 'utt/s', round(d['ms_per_step']*1e3,2), 'us/step', 'frac', round((d.get('roofline') or {}).get('frac',0),4), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print('$f failed', e)
PY
done
ls -la gpurun_out/r02_fused.ncu-rep
