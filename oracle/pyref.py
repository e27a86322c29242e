"""Pure-Python restatement of the reference's host functions.  TEST INFRASTRUCTURE ONLY.

Small-case oracle: each function states what the upstream function computes (file:line cited,
paths relative to the upstream repository) in independent code.  tests/golden/make_golden.py
checks every one of them against the real upstream functions (imported from /root/reference in
the build container) before writing the fixtures the GPU tests use.
"""
import math

import numpy as np


def edit_dist(s1, s2):
    """metrics.py:4-21.  s1 = reference (columns), s2 = hypothesis (rows); str or list[str].
    Returns (distance, len(s1)) like the reference (metrics.py:21)."""
    n_ref, n_hyp = len(s1), len(s2)
    prev = list(range(n_ref + 1))                      # row 0, metrics.py:13
    for i in range(1, n_hyp + 1):
        cur = [i] + [0] * n_ref                        # column 0, metrics.py:14
        for j in range(1, n_ref + 1):
            if s2[i - 1] == s1[j - 1]:
                cur[j] = prev[j - 1]                   # metrics.py:17-18
            else:
                cur[j] = 1 + min(cur[j - 1], prev[j - 1], prev[j])   # metrics.py:20
        prev = cur
    return int(prev[n_ref]), n_ref


def evaluate(s1, s2):
    """metrics.py:23-31: (CER, WER); the WER tokenisation is str.split(" ") (metrics.py:27-28).
    Empty reference -> ZeroDivisionError, as upstream."""
    d, n = edit_dist(s1, s2)
    cer = d / n
    d, n = edit_dist(s1.split(" "), s2.split(" "))
    return cer, d / n


def collapse_fn(preds):
    """CTCdecoder.py:119-131: drop every symbol equal to its predecessor.  No blank handling."""
    kept = [c for i, c in enumerate(preds) if i == 0 or c != preds[i - 1]]
    return "".join(kept)


def collapse_ids(ids, blank=0):
    """The CTC map B(): merge repeats on the frame path, then drop blanks (blank=0 as in
    CTCdecoder.py:41).  blank=None reproduces collapse_fn on ids."""
    out = []
    for i, c in enumerate(ids):
        if i > 0 and c == ids[i - 1]:
            continue
        if blank is not None and c == blank:
            continue
        out.append(int(c))
    return out


def reward_from_hyp(true_y, hyp, t):
    """policy_grad.py:10-15 applied to an already decoded+collapsed hypothesis `hyp`, with
    edit_dist(...)[0] where upstream subtracts the tuples (and raises TypeError)."""
    if t > 1:
        return -(edit_dist(true_y, hyp[:t + 1])[0] - edit_dist(true_y, hyp[:t])[0])
    if t == 1:
        return -(edit_dist(true_y, hyp[:t + 1])[0] - len(true_y))
    raise UnboundLocalError("r_t is unbound for t < 1 (policy_grad.py:10-16 has no else branch)")


def reward(true_y, pred_y, t, ind2char, ctc_decoder):
    """policy_grad.py:4-16 (intent): beam decode (beam_size=5, :6), ids -> chars (:7),
    collapse_fn (:8), then the incremental edit-distance reward."""
    ids, _ = ctc_decoder.decode(pred_y, beam_size=5)
    hyp = collapse_fn("".join(ind2char[i] for i in ids))
    return reward_from_hyp(true_y, hyp, t)


def nll_sum(inp, target, ignore_index=None):
    """loss.py:13-17 on numpy arrays: sum_i mean_b(-inp[i, b, target[b, i]]).  A falsy
    ignore_index (None or 0) ignores nothing (loss.py:9-12)."""
    inp = np.asarray(inp, np.float64)
    target = np.asarray(target)
    L, B, _ = inp.shape
    total = 0.0
    for i in range(L):
        picked = [-inp[i, b, target[b, i]] for b in range(B)
                  if not (ignore_index and target[b, i] == ignore_index)]
        total += sum(picked) / len(picked) if picked else float("nan")
    return total


def _lse(*xs):
    m = max(xs)
    if m == -math.inf:
        return -math.inf
    return m + math.log(sum(math.exp(x - m) for x in xs))


def prefix_beam_search(probs, beam_size=100, blank=0):
    """CTCdecoder.py:41-116 (Hannun's prefix beam search): probs [T,V] post-softmax.
    Returns (labels tuple, negative log-likelihood).  Candidate order and the stable sort on the
    total score follow the reference so ties resolve identically (:63-74 loop order, :110-113)."""
    T, V = probs.shape
    with np.errstate(divide="ignore"):
        logp = np.log(probs)                                                   # :55
    beam = [((), (0.0, -math.inf))]                                            # :60
    for t in range(T):
        nxt = {}
        order = []

        def get(pfx):
            if pfx not in nxt:
                nxt[pfx] = (-math.inf, -math.inf)
                order.append(pfx)
            return nxt[pfx]

        for s in range(V):                                                     # :66
            p = float(logp[t, s])
            for prefix, (p_b, p_nb) in beam:                                   # :72
                if s == blank:                                                 # :76-80
                    nb, nnb = get(prefix)
                    nxt[prefix] = (_lse(nb, p_b + p, p_nb + p), nnb)
                    continue
                last = prefix[-1] if prefix else None
                ext = prefix + (s,)
                nb, nnb = get(ext)
                if s != last:                                                  # :88-89
                    nnb = _lse(nnb, p_b + p, p_nb + p)
                else:                                                          # :90-94
                    nnb = _lse(nnb, p_b + p)
                nxt[ext] = (nb, nnb)
                if s == last:                                                  # :101-104
                    nb, nnb = get(prefix)
                    nxt[prefix] = (nb, _lse(nnb, p_nb + p))
        ranked = sorted(((k, nxt[k]) for k in order), key=lambda kv: _lse(*kv[1]), reverse=True)
        beam = ranked[:beam_size]                                              # :110-113
    best = beam[0]
    return best[0], -_lse(*best[1])
