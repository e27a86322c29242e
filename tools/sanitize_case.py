"""Small fused-step and stand-alone calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pgasr_b200 import functional as F
from tests.synth import make_batch
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(a).to(dev)
for (B, T, V, K, L, ragged) in ((3, 61, 30, 4, 9, True), (2, 900, 12, 3, 140, False), (2, 130, 30, 8, 40, True)):
    lg, tg, il, tl, uni = make_batch(B, T, V, K, L, seed=5, ragged=ragged)
    out = F.pg_ctc_step(t(lg), t(tg), t(il), t(tl), uniforms=t(uni), want=("rewards", "nll", "samples"))
    nll, g = F.ctc_loss_grad(t(lg), t(tg), t(il), t(tl))
    torch.cuda.synchronize()
    print("case", B, T, V, K, L, "loss", float(out["loss"]), "nll", nll.cpu().numpy()[:2])
p = np.random.default_rng(0).random((3, 40, 6)); p /= p.sum(-1, keepdims=True)
print(F.ctc_beam_search(t(p), None, beam_size=5)[1].cpu().numpy())
