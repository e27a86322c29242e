"""Throughput of the GPU prefix beam search next to the oracle restatement of upstream's decoder.
T=500, V=30 posteriors; beam_size=5 is what upstream's reward() and predict() use, 100 is decode()'s default."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pyref  # noqa: E402
from pgasr_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
T, V = 500, 30
out = []
for N, beam in ((148, 5), (592, 5), (148, 100)):
    z = rng.normal(size=(N, T, V)) * 2
    p = np.exp(z - z.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    pd = torch.from_numpy(p).to(dev)
    for _ in range(2):
        F.ctc_beam_search(pd, None, beam_size=beam)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        labels, label_len, nll = F.ctc_beam_search(pd, None, beam_size=beam)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    ref_labels, ref_nll = pyref.prefix_beam_search(p[0], beam_size=beam)
    cpu_s = time.perf_counter() - t0
    ok = tuple(labels[0, :int(label_len[0])].tolist()) == tuple(ref_labels) and abs(float(nll[0]) - ref_nll) < 1e-6
    out.append({"N": N, "T": T, "V": V, "beam": beam, "gpu_ms_per_launch": ms, "gpu_utt_per_s": N / ms * 1e3,
                "cpu_oracle_s_per_utt": cpu_s, "cpu_oracle_utt_per_s_1core": 1 / cpu_s, "matches_oracle": bool(ok)})
    print(json.dumps(out[-1]))
