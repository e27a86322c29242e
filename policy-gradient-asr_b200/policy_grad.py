"""Drop-in for upstream policy_grad.py: the per-position edit-distance reward, on the GPU.

reward(true_y, pred_y, t, ind2char, ctc_decoder) follows upstream policy_grad.py:4-16 with
edit_dist(...)[0] where upstream subtracts the (distance, length) tuples and raises TypeError:
   r_t = -(ED(y*, yhat[:t+1]) - ED(y*, yhat[:t]))   for t > 1
   r_1 = -(ED(y*, yhat[:2])   - len(y*))
One bit-parallel pass over the hypothesis yields ED(y*, yhat[:i]) for every i (the last column of the
DP table), so reward_all() returns every r_t from a single kernel launch.
"""
import torch

from . import functional as F
from .CTCdecoder import collapse_fn
from .metrics import _device, _encode_pair


def prefix_distances(true_y, hyp):
    """c[i] = ED(true_y, hyp[:i]) for i = 0..len(hyp), as a python list."""
    a, b, vocab = _encode_pair(true_y, hyp)
    if vocab > 256 or len(a) > 512:
        raise ValueError("prefix_distances: at most 256 distinct symbols and 512 reference symbols")
    dev = _device()
    ref = torch.tensor([a if a else [0]], dtype=torch.int32, device=dev)
    hyp_t = torch.tensor([b if b else [0]], dtype=torch.uint8, device=dev)
    _, col = F.edit_distance(hyp_t, torch.tensor([len(b)], dtype=torch.int32, device=dev), ref,
                             torch.tensor([len(a)], dtype=torch.int32, device=dev), rows_per_ref=1,
                             vocab=max(vocab, 1), last_col=True)
    return col[0, :len(b) + 1].cpu().tolist()


def reward_from_hyp(true_y, hyp, t, _col=None):
    col = _col if _col is not None else prefix_distances(true_y, hyp)
    n = len(col) - 1
    if t > 1:
        return -(col[min(t + 1, n)] - col[min(t, n)])
    if t == 1:
        return -(col[min(t + 1, n)] - len(true_y))
    raise UnboundLocalError("cannot access local variable 'r_t' where it is not associated with a value")


def reward_all(true_y, hyp):
    """[r_1, ..., r_len(hyp)] from one launch."""
    col = prefix_distances(true_y, hyp)
    return [reward_from_hyp(true_y, hyp, t, col) for t in range(1, len(hyp) + 1)]


def reward(true_y, pred_y, t, ind2char, ctc_decoder):
    """Upstream policy_grad.py:4-16."""
    pred_ids, _score = ctc_decoder.decode(pred_y, beam_size=5)
    hyp = "".join([ind2char[ind] for ind in pred_ids])
    hyp = collapse_fn(hyp)
    return reward_from_hyp(true_y, hyp, t)
