#!/bin/bash
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step']*1e3,2))"; }
echo "plain python N=1:"; run python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
echo "torchrun N=2:"; run $TR --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
echo "torchrun N=2, CUDA_DEVICE_MAX_CONNECTIONS=32:"; CUDA_DEVICE_MAX_CONNECTIONS=32 run $TR --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
echo "torchrun N=2, no NCCL (PGASR_BENCH_NO_DIST):"; PGASR_BENCH_NO_DIST=1 run $TR --master-port 29633 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
echo "torchrun N=2, OMP_NUM_THREADS=8:"; OMP_NUM_THREADS=8 run $TR --master-port 29634 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
echo "two plain python processes at once (GPU 0 and 1):"; (CUDA_VISIBLE_DEVICES=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > /tmp/b1.json 2>/dev/null &) ; run python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e; sleep 5; python -c "import json; d=json.loads(open('/tmp/b1.json').read().strip().splitlines()[-1]); print('gpu1', round(d['value']), round(d['ms_per_step']*1e3,2))"
