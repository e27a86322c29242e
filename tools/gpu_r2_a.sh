#!/bin/bash
# round 2, call A: tests + bench A/B of the launch path + batch sizes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_s20_$i.json 2> gpurun_out/r2a_bench_s20_$i.err; done
timeout 300 python bench.py > gpurun_out/r2a_bench_s400.json 2> gpurun_out/r2a_bench_s400.err
timeout 300 python bench.py --python-loop --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_pyloop.json 2> gpurun_out/r2a_bench_pyloop.err
timeout 300 python bench.py --python-loop --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_pyloop_s20.json 2> gpurun_out/r2a_bench_pyloop_s20.err
for b in 74 128 148 256; do timeout 300 python bench.py --batch $b --no-cpu-baseline > gpurun_out/r2a_bench_b$b.json 2> gpurun_out/r2a_bench_b$b.err; done
timeout 300 python bench.py --regime peaky --no-cpu-baseline > gpurun_out/r2a_bench_peaky.json 2> gpurun_out/r2a_bench_peaky.err
timeout 300 python bench.py --config 2 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
head -c 600 gpurun_out/r2a_bench_s20_1.json; echo
head -c 300 gpurun_out/r2a_bench_s400.json; echo
head -c 300 gpurun_out/r2a_bench_pyloop.json; echo
