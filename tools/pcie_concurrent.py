"""Host<->device copy bandwidth with every rank of the node copying at once (what bounds the e2e metric at N > 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_concurrent.py [out.json]

Each rank drives its own GPU: 3.84 MB copies (one step's logits / dlogits) from/to pinned host memory, H2D only, D2H
only and both directions at once, 300 copies per direction between barriers.  Rank 0 prints per-rank and aggregate
GB/s per direction, the CPU affinity of every rank and `nvidia-smi topo -m`."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 64 * 500 * 30
N = 300
h_in = [torch.randn(n).pin_memory() for _ in range(4)]
h_out = [torch.empty(n).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, device=dev) for _ in range(4)]
d_out = [torch.randn(n, device=dev) for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(mode):
    def go():
        for i in range(N):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)
    go()
    barrier()
    t0 = time.perf_counter()
    go()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return n * 4 * N / dt / 1e9          # GB/s per direction on this rank


res = {m: run(m) for m in ("h2d", "d2h", "both")}
aff = sorted(os.sched_getaffinity(0))
mine = {"rank": rank, "gbps": res, "cpus": len(aff), "cpu_first_last": [aff[0], aff[-1]]}
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, mine)
else:
    allr = [mine]
if rank == 0:
    out = {"n_gpus": world, "copy_mb": n * 4 / 1e6, "per_rank": allr,
           "aggregate_gbps_per_direction": {m: sum(r["gbps"][m] for r in allr) for m in ("h2d", "d2h", "both")},
           "host_cores": os.cpu_count()}
    try:
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception as e:
        out["topo"] = repr(e)
    txt = json.dumps(out)
    print(txt)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "a") as f:
            f.write(txt + "\n")
if world > 1:
    dist.destroy_process_group()
