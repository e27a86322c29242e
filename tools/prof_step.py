"""Small driver for ncu: a few PG+CTC steps at BASELINE.json configs[1] size (B=64,T=500,V=30,K=16,L=100)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pgasr_b200 import functional as F  # noqa: E402
from tests.synth import make_batch  # noqa: E402

B = int(os.environ.get("PROF_B", "64"))
steps = int(os.environ.get("PROF_STEPS", "4"))
dev = torch.device("cuda:0")
lg, tg, il, tl, _ = make_batch(B, 500, 30, 16, 100, seed=1)
t = lambda a: torch.from_numpy(a).to(dev)
lg, tg, il, tl = t(lg), t(tg), t(il), t(tl)
ws = None
modes = [(1.0, 1.0), (0.0, 1.0), (1.0, 0.0)] if os.environ.get("PROF_MODES") else [(1.0, 1.0)]
for wp, wc in modes:
    for i in range(steps):
        out = F.pg_ctc_step(lg, tg, il, tl, K=16, seed=i, workspace=ws, pg_weight=wp, ctc_weight=wc)
        ws = out["workspace"]
torch.cuda.synchronize()
print("loss", float(out["loss"]))
