#!/bin/bash
# multi-GPU round: tools/gpu_multi.sh N [with4]   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 $TR --master-port 29611 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 200 gpurun_out/r02_bench_n$N.err
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_s20_n$N.json 2> gpurun_out/r02_bench_s20_n$N.err
timeout 300 $TR --master-port 29613 tools/pcie_concurrent.py gpurun_out/r02_pcie_concurrent.jsonl > gpurun_out/pcie_n$N.log 2>&1; tail -c 300 gpurun_out/pcie_n$N.log
timeout 600 $TR --master-port 29614 bench.py --gpus $N --config 3 --steps 10 > gpurun_out/r02_bench_config3_n$N.json 2> gpurun_out/r02_bench_config3_n$N.err; tail -c 200 gpurun_out/r02_bench_config3_n$N.err
if [ "$2" = "with4" ]; then
  timeout 900 $TR --master-port 29615 bench.py --gpus $N --config 4 --steps 100 > gpurun_out/r02_bench_config4_n$N.json 2> gpurun_out/r02_bench_config4_n$N.err; tail -c 200 gpurun_out/r02_bench_config4_n$N.err
fi
python - <<PY
import json
for f in ("r02_bench_n$N", "r02_bench_s20_n$N", "r02_bench_config3_n$N", "r02_bench_config4_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, round(d["value"]), "utt/s", round(d["ms_per_step"], 4), "ms/step  e2e", e.get("value"), e.get("ms_per_step"), (d.get("breakdown") or {}).get("allreduce_exposed_ms"))
    except Exception as ex:
        print(f, "n/a", ex)
PY
