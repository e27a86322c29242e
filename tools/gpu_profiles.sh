#!/bin/bash
# round-2 profile set: launch list of the bench command, --set full of the fused kernel and of the stand-alone kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "edit_distance or collapse or beam" > gpurun_out/p_pytest.log 2>&1; tail -2 gpurun_out/p_pytest.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/p_ncu_launches.log 2>&1
timeout 300 python tools/prof_step.py > gpurun_out/p_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 \
    -k regex:pg_ctc_fused -s 2 -c 1 -f -o gpurun_out/r02_fused python tools/prof_step.py > gpurun_out/p_ncu_fused.log 2>&1
timeout 300 python tools/prof_standalone.py > gpurun_out/p_standalone.log 2>&1 || { tail -5 gpurun_out/p_standalone.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'softmax_sample|collapse_u8|myers_u8|wavefront_i32|pg_grad_kernel|pg_advantages|ctc_beam' \
    -s 9 -c 9 -f -o gpurun_out/r02_standalone python tools/prof_standalone.py > gpurun_out/p_ncu_standalone.log 2>&1
tail -2 gpurun_out/p_ncu_standalone.log
ls -la gpurun_out/*.ncu-rep
