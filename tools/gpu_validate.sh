#!/bin/bash
# full suite + long fuzz + reward-to-go bench line
TAG=${1:-v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 1200 python tools/fuzz_step.py 350 41 > gpurun_out/${TAG}_fuzz350.log 2>&1; tail -2 gpurun_out/${TAG}_fuzz350.log; grep -c "^skip" gpurun_out/${TAG}_fuzz350.log
timeout 300 python bench.py --reward ed_to_go --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_togo.json 2> gpurun_out/${TAG}_bench_togo.err; head -c 250 gpurun_out/${TAG}_bench_togo.json; echo
