"""Generate tests/golden/reference_vectors.json from the REAL upstream functions.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
It imports metrics.py / CTCdecoder.py / loss.py from /root/reference, evaluates them on the
known-answer inputs of SURVEY.md section 8(c) plus seeded random cases, asserts that both oracle
restatements (oracle/pyref.py and oracle/pgasr_oracle.c) reproduce every value, and writes the
inputs and upstream outputs as the fixture the CPU and GPU tests replay.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import metrics as ref_metrics            # noqa: E402  (upstream)
import CTCdecoder as ref_ctc             # noqa: E402  (upstream)
import loss as ref_loss                  # noqa: E402  (upstream)
import torch                             # noqa: E402

from oracle import cport, pyref          # noqa: E402


def enc(s):
    return [ord(c) for c in s]


def main():
    out = {"generator": "tests/golden/make_golden.py", "upstream": "ana-kuznetsova/Policy-Gradient-ASR"}

    # ---- edit_dist: known answers + seeded random -----------------------------
    ed = []
    for a, b in [("kitten", "sitting"), ("", "abc"), ("abc", ""), ("", ""), ("a", "a"),
                 ("hello world", "helo wurld!"), ("the cat sat", "the bat sat on")]:
        ed.append({"kind": "str", "ref": a, "hyp": b, "out": list(ref_metrics.edit_dist(a, b))})
    for a, b in [("a b c", "a x c d"), ("the cat sat", "the bat sat on"), ("a  b", " a b ")]:
        ed.append({"kind": "words", "ref": a, "hyp": b,
                   "out": list(ref_metrics.edit_dist(a.split(" "), b.split(" ")))})
    rng = np.random.default_rng(20261018)
    for (lr, lh, v) in [(5, 7, 4), (100, 100, 30), (200, 137, 30), (400, 400, 30), (1, 50, 30), (50, 1, 30)]:
        a = rng.integers(1, v, lr).tolist()
        b = rng.integers(1, v, lh).tolist()
        ed.append({"kind": "ids", "ref": a, "hyp": b, "out": list(ref_metrics.edit_dist(a, b))})
    coll = []
    for (t, v) in [(20, 3), (500, 30)]:
        path = rng.integers(0, v, t).tolist()
        s = "".join(chr(48 + x) for x in path)
        c = ref_ctc.collapse_fn(s)
        coll.append({"path": path, "collapsed": [ord(ch) - 48 for ch in c]})
    # more random edit distances incl. tiny alphabets (many matches) and 128/129 word boundaries
    rng2 = np.random.default_rng(7)
    for (lr, lh, v) in [(31, 33, 3), (32, 32, 2), (33, 64, 5), (64, 65, 4), (127, 130, 6),
                        (128, 128, 3), (129, 127, 30), (255, 300, 8), (256, 20, 30), (0, 17, 5), (17, 0, 5)]:
        a = rng2.integers(1, v, lr).tolist()
        b = rng2.integers(1, v, lh).tolist()
        ed.append({"kind": "ids", "ref": a, "hyp": b, "out": list(ref_metrics.edit_dist(a, b))})
    for e in ed:
        a, b = (e["ref"], e["hyp"])
        if e["kind"] == "words":
            a, b = a.split(" "), b.split(" ")
        assert list(pyref.edit_dist(a, b)) == e["out"], e
        if e["kind"] == "ids":
            assert cport.edit_distance(a, b) == e["out"][0], e
        elif e["kind"] == "str":
            assert cport.edit_distance(enc(a), enc(b)) == e["out"][0], e
    out["edit_dist"] = ed

    # ---- evaluate --------------------------------------------------------------------------
    ev = []
    for a, b in [("the cat sat", "the bat sat on"), ("hello world", "helo wurld!"), ("a", "b"),
                 ("ab cd", "ab  cd")]:
        r = ref_metrics.evaluate(a, b)
        assert pyref.evaluate(a, b) == r
        ev.append({"ref": a, "hyp": b, "out": [r[0], r[1]]})
    out["evaluate"] = ev

    # ---- collapse_fn -------------------------------------------------------------------------
    cf = []
    for s in ["", "a", "aabbcc", "aa_bb__a", "hello  world", "abab", "aaa", "__a__", "abba"]:
        c = ref_ctc.collapse_fn(s)
        assert pyref.collapse_fn(s) == c
        assert "".join(chr(x) for x in cport.collapse(enc(s), blank=-1)) == c
        cf.append({"in": s, "out": c})
    out["collapse_fn"] = cf
    for c in coll:
        assert cport.collapse(c["path"], blank=-1).tolist() == c["collapsed"]
        assert pyref.collapse_ids(c["path"], blank=None) == c["collapsed"]
        c["collapsed_no_blank"] = [x for x in c["collapsed"] if x != 0]
        assert cport.collapse(c["path"], blank=0).tolist() == c["collapsed_no_blank"]
    out["collapse_paths"] = coll

    # ---- per-position reward (upstream raises; pin the intent via upstream edit_dist) ---------
    rw = []
    for y, h in [("hello world", "helo wurld!"), ("abc", "abc"), ("abc", ""), ("kitten", "sitting")]:
        rs = []
        for t in range(1, len(h) + 3):
            if t > 1:
                r = -(ref_metrics.edit_dist(y, h[:t + 1])[0] - ref_metrics.edit_dist(y, h[:t])[0])
            else:
                r = -(ref_metrics.edit_dist(y, h[:t + 1])[0] - len(y))
            assert pyref.reward_from_hyp(y, h, t) == r
            rs.append(r)
        cr = cport.reward_positions(enc(y), enc(h), len(h) + 2)
        assert cr[1:].tolist() == rs, (cr, rs)
        rw.append({"true_y": y, "hyp": h, "r": rs})
    out["reward_positions"] = rw

    # ---- customNLLLoss -------------------------------------------------------------------------
    torch.manual_seed(3)
    L, B, V = 5, 4, 7
    inp = torch.log_softmax(torch.randn(L, B, V), -1)
    tgt = torch.randint(0, V, (B, L))
    nl = []
    for ign in [None, 0, 2]:
        val = float(ref_loss.customNLLLoss(ignore_index=ign)(inp, tgt))
        mine = pyref.nll_sum(inp.numpy(), tgt.numpy(), ign)
        assert abs(mine - val) < 1e-5, (ign, mine, val)
        cval = cport.nll_sum(inp.numpy(), tgt.numpy(), ign if ign else -1)
        assert abs(cval - val) < 1e-5, (ign, cval, val)
        nl.append({"ignore_index": ign, "out": val})
    out["nll"] = {"inp": inp.numpy().round(6).tolist(), "target": tgt.tolist(), "cases": nl}
    # re-evaluate on the rounded inputs so the fixture is self-consistent
    inp_r = torch.tensor(out["nll"]["inp"], dtype=torch.float32)
    for c in nl:
        c["out"] = float(ref_loss.customNLLLoss(ignore_index=c["ignore_index"])(inp_r, tgt))

    # ---- prefix beam search ------------------------------------------------------------------
    bs = []
    rng3 = np.random.default_rng(11)
    dec = ref_ctc.CTCDecoder(alphabet=None)
    for (T, V, beam) in [(6, 3, 2), (12, 4, 5), (30, 6, 5), (25, 5, 100)]:
        z = rng3.normal(size=(T, V)) * 2
        p = np.exp(z - z.max(1, keepdims=True))
        p = (p / p.sum(1, keepdims=True)).round(6)
        labels, nll = dec.decode(p, beam_size=beam)
        ml, mn = pyref.prefix_beam_search(p, beam_size=beam)
        assert tuple(ml) == tuple(labels) and abs(mn - nll) < 1e-9, (labels, ml, nll, mn)
        bs.append({"probs": p.tolist(), "beam": beam, "labels": list(map(int, labels)), "nll": float(nll)})
    out["beam_search"] = bs

    path = os.path.join(HERE, "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
