"""e2e pipeline throughput: full step vs copies only (weights 0), several depths, repeated (the number is noisy)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pgasr_b200
from tests.synth import make_batch
B,T,V,K,L=64,500,30,16,100
pool=[]
for i in range(8):
    lg,tg,il,tl,_=make_batch(B,T,V,K,L,seed=i)
    pool.append(tuple(torch.from_numpy(a).pin_memory() for a in (lg,tg,il,tl)))
N=int(os.environ.get("N","1500"))
for name,kw in (("copies only (weights 0)",dict(pg_weight=0.0,ctc_weight=0.0)),("full",dict())):
    for depth in (2,3,4,6):
        pipe=pgasr_b200.HostPipeline(B,T,V,K,L,depth=depth,**kw)
        outs=[pipe.output_buffers() for _ in range(depth)]
        res=[]
        for rep in range(4):
            for i in range(20): pipe.submit(*pool[i%8],out=outs[i%depth],seed=i)
            pipe.wait(); torch.cuda.synchronize()
            t0=time.perf_counter()
            for i in range(N): pipe.submit(*pool[i%8],out=outs[i%depth],seed=i)
            pipe.wait(); dt=time.perf_counter()-t0
            res.append(dt/N*1e6)
        print(f"{name:26s} depth {depth}: " + "  ".join(f"{r:6.1f}" for r in res) + " us/step")
        pipe.close()
