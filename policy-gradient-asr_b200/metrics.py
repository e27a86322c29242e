"""Drop-in for upstream metrics.py: edit_dist / evaluate / save_predictions, computed on the GPU.

Signatures, return conventions and error behaviour follow upstream (metrics.py:4-37):
  edit_dist(s1, s2) -> (distance, len(s1)); s1 is the reference; str (CER) or list[str] (WER)
  evaluate(s1, s2)  -> (cer, wer); ZeroDivisionError on an empty reference, like upstream
"""
import os

import torch

from . import functional as F


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pgasr_b200.metrics needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _encode_pair(s1, s2):
    """Map the symbols of both sequences to dense int ids (equality is all the distance looks at)."""
    table = {}
    a = [table.setdefault(x, len(table)) for x in s1]
    b = [table.setdefault(x, len(table)) for x in s2]
    return a, b, len(table)


def edit_dist_batch(refs, hyps):
    """Distances for a list of (reference, hypothesis) pairs in one launch.  Returns a list of ints."""
    if len(refs) != len(hyps):
        raise ValueError("refs and hyps must pair up")
    n = len(refs)
    if n == 0:
        return []
    enc = [_encode_pair(r, h) for r, h in zip(refs, hyps)]
    lr = max(1, max(len(e[0]) for e in enc))
    lh = max(1, max(len(e[1]) for e in enc))
    vocab = max(e[2] for e in enc)
    dev = _device()
    ref_t = torch.zeros((n, lr), dtype=torch.int32)
    hyp_t = torch.zeros((n, lh), dtype=torch.int32)
    for i, (a, b, _) in enumerate(enc):
        ref_t[i, :len(a)] = torch.tensor(a, dtype=torch.int32)
        hyp_t[i, :len(b)] = torch.tensor(b, dtype=torch.int32)
    ref_len = torch.tensor([len(e[0]) for e in enc], dtype=torch.int32)
    hyp_len = torch.tensor([len(e[1]) for e in enc], dtype=torch.int32)
    if vocab <= 256 and lr <= 512:
        d = F.edit_distance(hyp_t.to(torch.uint8).to(dev), hyp_len.to(dev), ref_t.to(dev), ref_len.to(dev),
                            rows_per_ref=1, vocab=max(vocab, 1))
    else:
        d = F.edit_distance_tokens(hyp_t.to(dev), hyp_len.to(dev), ref_t.to(dev), ref_len.to(dev))
    return [int(x) for x in d.cpu().tolist()]


def edit_dist(s1, s2):
    """Upstream metrics.py:4-21.  s1: reference, s2: prediction; str or list[str]."""
    return edit_dist_batch([s1], [s2])[0], len(s1)


def evaluate(s1, s2):
    """Upstream metrics.py:23-31: character error rate, then word error rate on str.split(" ")."""
    w1, w2 = s1.split(" "), s2.split(" ")
    d_char, d_word = edit_dist_batch([s1, w1], [s2, w2])
    cer = d_char / len(s1)
    wer = d_word / len(w1)
    return cer, wer


def evaluate_batch(targets, predictions):
    """evaluate() for a whole list of (target, prediction) string pairs (upstream's predict() loop calls it per
    utterance, model.py:321-337): two distances per pair, all pairs in one launch per kernel.
    Returns a list of (cer, wer); an empty target raises ZeroDivisionError like upstream."""
    if len(targets) != len(predictions):
        raise ValueError("targets and predictions must pair up")
    refs, hyps = [], []
    for s1, s2 in zip(targets, predictions):
        refs += [s1, s1.split(" ")]
        hyps += [s2, s2.split(" ")]
    d = edit_dist_batch(refs, hyps)
    out = []
    for i, s1 in enumerate(targets):
        out.append((d[2 * i] / len(s1), d[2 * i + 1] / len(s1.split(" "))))
    return out


def save_predictions(target, predicted, model_path):
    """Upstream metrics.py:33-37: one `target|prediction` line per utterance in predicted.txt."""
    path = os.path.join(model_path, "predicted.txt")
    with open(path, "w") as fo:
        for i, tgt in enumerate(target):
            fo.write("|".join((tgt, predicted[i])) + "\n")
