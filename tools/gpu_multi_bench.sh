#!/bin/bash
# bench only (400 steps and the driver's 20 steps) under torchrun: tools/gpu_multi_bench.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29621 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 200 gpurun_out/r02_bench_n$N.err
timeout 600 $TR --master-port 29622 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_s20_n$N.json 2> gpurun_out/r02_bench_s20_n$N.err
python - <<PY
import json
for f in ("r02_bench_n$N", "r02_bench_s20_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, round(d["value"]), "utt/s", round(d["ms_per_step"]*1e3, 2), "us/step  e2e", round(e.get("value") or 0), round((e.get("ms_per_step") or 0)*1e3,1))
    except Exception as ex:
        print(f, "n/a", ex)
PY
