"""The inner loop of upstream predict() (model.py:321-334) as three batched GPU calls.

Per utterance upstream does: probs = exp(log_probs[:pad_ind]) -> CTCDecoder.decode(beam_size=5) -> ids to chars ->
collapse_fn -> evaluate(target, seq).  `pad_ind` is the TRANSCRIPT length, int(sum(tmask[i])) (model.py:322-323) --
upstream cuts the frames with it; that quirk, and the second repeat-merge after the decoder already merged repeats
(collapse_fn turns "hello" into "helo"), are kept so that the numbers match upstream's.
"""
import numpy as np

from . import metrics
from .CTCdecoder import CTCDecoder, collapse_batch


def decode_and_score(preds, t, tmask, ind2char, beam_size=5, blank=0, ctc_decoder=None):
    """preds [B,T,V] log-probabilities (numpy, as after .detach().cpu().numpy(), model.py:317), t [B,L] label ids,
    tmask [B,L]; ind2char maps ids to characters.  Returns (targets, predicted, cers, wers): the lists upstream appends
    to / accumulates (model.py:326-334), one entry per utterance."""
    preds = np.asarray(preds)
    t = np.asarray(t)
    tmask = np.asarray(tmask)
    dec = ctc_decoder if ctc_decoder is not None else CTCDecoder(alphabet=None)
    pads = [int(np.sum(tmask[i])) for i in range(len(preds))]
    probs = [np.exp(preds[i][:pads[i]]) for i in range(len(preds))]            # model.py:322-323
    seqs = dec.decode_batch(probs, beam_size=beam_size, blank=blank)           # :324
    hyps = ["".join(ind2char[ind] for ind in seq) for seq, _ in seqs]          # :325
    predicted = collapse_batch(hyps)                                           # :326
    targets = ["".join(ind2char[ind] for ind in t[i][:pads[i]]) for i in range(len(preds))]   # :327-329
    scores = metrics.evaluate_batch(targets, predicted)                        # :332
    return targets, predicted, [c for c, _ in scores], [w for _, w in scores]
