"""Randomised check of the GPU prefix beam search against the oracle restatement of upstream's decoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pyref
from pgasr_b200 import functional as F
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
bad = 0
for case in range(n):
    T = int(rng.integers(1, 90)); V = int(rng.choice([2, 3, 5, 8, 30, 64])); beam = int(rng.choice([1, 2, 5, 16, 100, 128]))
    N = 4
    scale = float(rng.choice([0.5, 2.0, 6.0]))
    z = rng.normal(size=(N, T, V)) * scale
    p = np.exp(z - z.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    if rng.integers(0, 3) == 0:
        p = p.round(int(rng.choice([1, 2, 4])))                    # exact zeros and exact ties
    lens = rng.integers(1, T + 1, size=N).astype(np.int32)
    labels, label_len, nll = F.ctc_beam_search(torch.from_numpy(p).to(dev), torch.from_numpy(lens).to(dev), beam_size=beam)
    labels, label_len, nll = labels.cpu().numpy(), label_len.cpu().numpy(), nll.cpu().numpy()
    for i in range(N):
        with np.errstate(divide="ignore"):
            rl, rn = pyref.prefix_beam_search(p[i, :lens[i]], beam_size=beam)
        same = tuple(labels[i, :label_len[i]]) == tuple(rl)
        close = (np.isinf(rn) and np.isinf(nll[i])) or abs(nll[i] - rn) <= 1e-9 * max(1.0, abs(rn))
        if not (same and close):
            bad += 1
            print(f"MISMATCH case {case} utt {i}: T={lens[i]} V={V} beam={beam} scale={scale} gpu {tuple(labels[i,:label_len[i]])} {nll[i]} ref {tuple(rl)} {rn}")
print(f"{n * 4} decodes, {bad} mismatches")
sys.exit(1 if bad else 0)
