"""Drop-in for upstream CTCdecoder.py: collapse_fn (GPU) and CTCDecoder (prefix beam search).

collapse_fn(preds: str) -> str keeps a character iff it differs from its predecessor
(upstream CTCdecoder.py:119-131); it does not remove blanks, exactly like upstream.
CTCDecoder.decode is the eval-time prefix beam search (upstream CTCdecoder.py:41-116); it is outside the
training hot path and runs on the host.
"""
import math

import numpy as np
import torch

from . import functional as F

NEG_INF = -float("inf")


def collapse_batch(strings):
    """collapse_fn over a list of strings in one launch."""
    if not strings:
        return []
    for s in strings:
        if not isinstance(s, str):
            raise TypeError("can only concatenate str (not \"%s\") to str" % type(s).__name__)
    alphabet = sorted(set("".join(strings)))
    if len(alphabet) > 256:
        raise ValueError("collapse_fn: more than 256 distinct characters in one batch")
    if not torch.cuda.is_available():
        raise RuntimeError("pgasr_b200.CTCdecoder.collapse_fn needs a CUDA device (no CPU fallback)")
    to_id = {c: i for i, c in enumerate(alphabet)}
    T = max(1, max(len(s) for s in strings))
    rows = np.zeros((len(strings), T), np.uint8)
    for i, s in enumerate(strings):
        rows[i, :len(s)] = [to_id[c] for c in s]
    dev = torch.device("cuda", torch.cuda.current_device())
    lens = torch.tensor([len(s) for s in strings], dtype=torch.int32, device=dev)
    out, out_len = F.collapse(torch.from_numpy(rows).to(dev), lens, rows_per_len=1, blank=None)
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    return ["".join(alphabet[x] for x in out[i, :out_len[i]]) for i in range(len(strings))]


def collapse_fn(preds):
    """Upstream CTCdecoder.py:119-131."""
    return collapse_batch([preds])[0]


def _lse(*xs):
    m = max(xs)
    if m == NEG_INF:
        return NEG_INF
    return m + math.log(sum(math.exp(x - m) for x in xs))


class CTCDecoder:
    """Prefix beam search with the upstream constructor and decode() signature
    (CTCdecoder.py:23-25, 41-116): decode(probs[T,V] post-softmax, beam_size=100, blank=0)
    -> (labels tuple, negative log-likelihood)."""

    def __init__(self, alphabet):
        self.alphabet = alphabet
        self.NEG_INF = NEG_INF

    def make_new_beam(self):
        return {}

    def logsumexp(self, *args):
        return _lse(*args)

    def decode(self, probs, beam_size=100, blank=0):
        T, V = probs.shape
        with np.errstate(divide="ignore"):
            lp = np.log(probs)
        beam = [((), (0.0, NEG_INF))]
        for t in range(T):
            cand = {}        # insertion-ordered: ties in the sort below resolve as upstream's do

            def slot(prefix):
                if prefix not in cand:
                    cand[prefix] = [NEG_INF, NEG_INF]
                return cand[prefix]

            for s in range(V):
                p = float(lp[t, s])
                for prefix, (p_b, p_nb) in beam:
                    if s == blank:
                        e = slot(prefix)
                        e[0] = _lse(e[0], p_b + p, p_nb + p)
                        continue
                    last = prefix[-1] if prefix else None
                    e = slot(prefix + (s,))
                    e[1] = _lse(e[1], p_b + p, p_nb + p) if s != last else _lse(e[1], p_b + p)
                    if s == last:
                        e = slot(prefix)
                        e[1] = _lse(e[1], p_nb + p)
            ranked = sorted(cand.items(), key=lambda kv: _lse(*kv[1]), reverse=True)
            beam = [(k, (v[0], v[1])) for k, v in ranked[:beam_size]]
        labels, (p_b, p_nb) = beam[0]
        return labels, -_lse(p_b, p_nb)
