"""Driver for ncu: every stand-alone kernel once warm, then once more, at BASELINE.json configs[1] size
(B=64, T=500, V=30, K=16, L=100): softmax_sample, collapse_u8, myers_u8 (split and last-column flavours), wavefront_i32,
pg_advantages, pg_grad, ctc_beam (16 utterances, beam 100)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pgasr_b200 import functional as F  # noqa: E402
from tests.synth import make_batch  # noqa: E402

B, T, V, K, L = 64, 500, 30, 16, 100
dev = torch.device("cuda:0")
lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=1)
t = lambda a: torch.from_numpy(a).to(dev)
lg, tg, il, tl = t(lg), t(tg), t(il), t(tl)
for rep in range(2):
    smp, logp, probs = F.softmax_sample(lg, il, K=K, seed=1, return_probs=True)
    hyp, hl = F.collapse(smp, il, blank=0)
    dist = F.edit_distance(hyp, hl.reshape(-1), tg, tl, rows_per_ref=K, vocab=V)
    dist2, col = F.edit_distance(hyp, hl.reshape(-1), tg, tl, rows_per_ref=K, vocab=V, last_col=True)
    assert torch.equal(dist, dist2)
    rew, adv, terms = F.pg_advantages(dist, tl, logp, Lmax=L)
    g = F.pg_grad(smp, adv, il, V=V, scale=1.0 / (B * K))
    # word-level distances: int32 tokens, 64 pairs of ~300 x 300
    rng = np.random.default_rng(3)
    hw = t(rng.integers(0, 5000, (64, 300)).astype(np.int32))
    rw = t(rng.integers(0, 5000, (64, 300)).astype(np.int32))
    dw = F.edit_distance_tokens(hw, None, rw, None)
    lab, lab_len, nll = F.ctc_beam_search(probs[:16].double().contiguous(), il[:16].contiguous(), beam_size=100)
torch.cuda.synchronize()
print("ok", int(dist.sum()), int(dw.sum()), float(nll.sum()))
