"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the upstream golden vectors.

Bars: samples / collapsed hypotheses / edit distances / rewards bit-exact; CTC loss, CTC gradient, sequence
log-probabilities and the policy gradient within 1e-4 relative (gradients: max abs error over the tensor
relative to its max abs value), fp32, as BASELINE.json's north_star states.
"""
import math

import numpy as np
import pytest
import torch

from oracle import cport
from tests.synth import make_batch

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def dev_t(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=device, dtype=dtype) if dtype else t.to(device)


def rel_err(got, want):
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    return np.abs(got - want).max() / scale


def grad_close(got, want, rtol=RTOL):
    """Element-wise bar next to the max-norm one: |got - want| <= 1e-4 |want| + atol for EVERY entry, with
    atol = max(1e-7, 1e-5 max|want|): 1e-7 absolute at the bench shape (rows are scaled by 1/B = 1/64, mean baseline),
    and never below ~100 fp32 ulps of the largest entry: dlogits is fp32 and an entry is a difference of O(max)
    terms -- without a baseline the REINFORCE row is p * sum_k A_k minus K scatter terms of the same size, each
    rounded in fp32 (measured 1.6e-6 of the maximum at K = 33), so an entry cannot be closer to the fp64 oracle
    than that.  Still ten times tighter than the max-norm bar, and applied to every entry."""
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    atol = max(1e-7, 1e-5 * float(np.abs(want).max()))
    bad = np.abs(got - want) > rtol * np.abs(want) + atol
    return not bad.any()


def enc(s):
    return [ord(c) for c in s]


# ---------------------------------------------------------------- a6 sampler
@pytest.mark.parametrize("B,T,V,K,ragged", [(3, 37, 30, 4, True), (2, 300, 5, 16, False), (64, 500, 30, 16, False),
                                            (2, 700, 32, 7, True), (1, 1, 2, 1, False), (3, 90, 33, 5, True),
                                            (2, 200, 64, 16, False), (2, 64, 47, 3, True)])
def test_sampler_bit_exact_injected_uniforms(cuda, B, T, V, K, ragged):
    from pgasr_b200 import functional as F
    logits, _, in_len, _, uni = make_batch(B, T, V, K, max(T // 5, 1), seed=B + T, ragged=ragged)
    s_ref, lp_ref = cport.softmax_sample(logits, in_len, uni)
    s, lp, probs = F.softmax_sample(dev_t(logits, cuda), dev_t(in_len, cuda), uniforms=dev_t(uni, cuda), return_probs=True)
    assert np.array_equal(s.cpu().numpy(), s_ref)
    assert rel_err(lp.cpu().numpy(), lp_ref) < RTOL
    p_ref = torch.softmax(torch.tensor(logits, dtype=torch.float64), -1).numpy()
    for b in range(B):
        p_ref[b, in_len[b]:] = 0
    assert np.abs(probs.cpu().numpy() - p_ref).max() < 1e-6


def test_sampler_bit_exact_philox(cuda):
    from pgasr_b200 import functional as F
    logits, _, in_len, _, _ = make_batch(5, 211, 30, 16, 10, seed=9, ragged=True)
    for seed, K in [(0x5EED, 16), (2**63 + 12345, 6)]:
        s_ref, lp_ref = cport.softmax_sample(logits, in_len, None, seed=seed, K=K)
        s, lp = F.softmax_sample(dev_t(logits, cuda), dev_t(in_len, cuda), K=K, seed=seed)
        assert np.array_equal(s.cpu().numpy(), s_ref)
        assert rel_err(lp.cpu().numpy(), lp_ref) < RTOL


def test_sampler_extreme_logits(cuda):
    from pgasr_b200 import functional as F
    rng = np.random.default_rng(0)
    logits = (rng.standard_normal((2, 64, 30)) * 40).astype(np.float32)      # exp underflow below -87
    logits[0, 0, :] = 0.0
    logits[0, 1, :] = -1e4
    logits[0, 1, 7] = 50.0
    uni = np.minimum(rng.random((2, 8, 64), dtype=np.float32), np.float32(1 - 2**-24))
    uni[0, 0, :] = 0.0
    uni[0, 1, :] = np.float32(1 - 2**-24)
    s_ref, _ = cport.softmax_sample(logits, None, uni)
    s, _ = F.softmax_sample(dev_t(logits, cuda), uniforms=dev_t(uni, cuda))
    assert np.array_equal(s.cpu().numpy(), s_ref)
    assert (s_ref[0, :, 1] == 7).all()


# ---------------------------------------------------------------- a3 collapse
def test_collapse_golden_and_random(cuda, golden):
    from pgasr_b200 import functional as F
    for e in golden["collapse_paths"]:
        x = dev_t(np.array([e["path"]], np.uint8), cuda)
        out, n = F.collapse(x, blank=None)
        assert out[0, :int(n[0])].cpu().tolist() == e["collapsed"]
        out, n = F.collapse(x, blank=0)
        assert out[0, :int(n[0])].cpu().tolist() == e["collapsed_no_blank"]
    rng = np.random.default_rng(3)
    B, K, T = 7, 5, 333
    s = rng.integers(0, 4, (B, K, T)).astype(np.uint8)
    in_len = rng.integers(0, T + 1, B).astype(np.int32)
    in_len[0], in_len[1] = 0, T
    out, n = F.collapse(dev_t(s, cuda), dev_t(in_len, cuda), blank=0)
    out, n = out.cpu().numpy(), n.cpu().numpy()
    for b in range(B):
        for k in range(K):
            want = cport.collapse(s[b, k, :in_len[b]].astype(np.int32), blank=0)
            assert n[b, k] == len(want)
            assert np.array_equal(out[b, k, :len(want)], want.astype(np.uint8))
            assert (out[b, k, len(want):] == 0).all()


def test_dropin_collapse_fn(cuda, golden):
    import pgasr_b200
    for e in golden["collapse_fn"]:
        assert pgasr_b200.CTCdecoder.collapse_fn(e["in"]) == e["out"]
    with pytest.raises(TypeError):
        pgasr_b200.CTCdecoder.collapse_fn(["a", "a"])


# ---------------------------------------------------------------- a1 / a2 / a4 edit distance
def test_dropin_edit_dist_golden(cuda, golden):
    import pgasr_b200
    for e in golden["edit_dist"]:
        a, b = e["ref"], e["hyp"]
        if e["kind"] == "words":
            a, b = a.split(" "), b.split(" ")
        assert list(pgasr_b200.metrics.edit_dist(a, b)) == e["out"], e
    for e in golden["evaluate"]:
        assert list(pgasr_b200.metrics.evaluate(e["ref"], e["hyp"])) == e["out"]
    with pytest.raises(ZeroDivisionError):
        pgasr_b200.metrics.evaluate("", "abc")
    assert pgasr_b200.metrics.edit_dist("", "") == (0, 0)


def test_dropin_reward_golden(cuda, golden):
    import pgasr_b200
    for e in golden["reward_positions"]:
        y, h = e["true_y"], e["hyp"]
        got = [pgasr_b200.policy_grad.reward_from_hyp(y, h, t) for t in range(1, len(h) + 3)]
        assert got == e["r"]
    y, h = "hello world", "helo wurld!"
    r = pgasr_b200.policy_grad.reward_all(y, h)
    assert r[:10] == [2, 1, 1, 1, 1, 0, 1, 1, 1, -1] and sum(r) == 8
    with pytest.raises(UnboundLocalError):
        pgasr_b200.policy_grad.reward_from_hyp(y, h, 0)

    class FakeDecoder:                      # reward() through the upstream call shape
        def decode(self, probs, beam_size=100, blank=0):
            assert beam_size == 5
            return tuple(enc("heello")), 0.0
    ind2char = {i: chr(i) for i in range(128)}
    got = pgasr_b200.policy_grad.reward("hello", np.zeros((3, 3)), 2, ind2char, FakeDecoder())
    assert got == -(cport.edit_distance(enc("hello"), enc("hel")) - cport.edit_distance(enc("hello"), enc("he")))


@pytest.mark.parametrize("Lr,Th,V,rows", [(1, 50, 30, 3), (31, 40, 3, 4), (32, 33, 2, 16), (33, 64, 5, 1), (100, 500, 30, 16),
                                          (128, 200, 4, 2), (129, 130, 30, 5), (256, 300, 8, 2), (400, 450, 30, 3), (512, 64, 6, 2)])
def test_edit_distance_bit_exact_random(cuda, Lr, Th, V, rows):
    from pgasr_b200 import functional as F
    rng = np.random.default_rng(Lr * 1000 + Th)
    G = 6
    refs = rng.integers(1, V, (G, Lr)).astype(np.int32)
    ref_len = rng.integers(0, Lr + 1, G).astype(np.int32)
    ref_len[0], ref_len[1] = Lr, 0
    hyps = rng.integers(1, V, (G * rows, Th)).astype(np.uint8)
    hyp_len = rng.integers(0, Th + 1, G * rows).astype(np.int32)
    hyp_len[0], hyp_len[1] = Th, 0
    d, col = F.edit_distance(dev_t(hyps, cuda), dev_t(hyp_len, cuda), dev_t(refs, cuda), dev_t(ref_len, cuda),
                             rows_per_ref=rows, vocab=V, last_col=True)
    d2 = F.edit_distance(dev_t(hyps, cuda), dev_t(hyp_len, cuda), dev_t(refs, cuda), dev_t(ref_len, cuda),
                         rows_per_ref=rows, vocab=V)
    dw = F.edit_distance_tokens(dev_t(hyps.astype(np.int32), cuda), dev_t(hyp_len, cuda), dev_t(refs, cuda),
                                dev_t(ref_len, cuda), rows_per_ref=rows)
    d, col, d2, dw = d.cpu().numpy(), col.cpu().numpy(), d2.cpu().numpy(), dw.cpu().numpy()
    for r in range(G * rows):
        g = r // rows
        want, wcol = cport.edit_distance(refs[g, :ref_len[g]], hyps[r, :hyp_len[r]].astype(np.int32), last_col=True)
        assert d[r] == want and d2[r] == want and dw[r] == want, (r, d[r], d2[r], dw[r], want)
        assert np.array_equal(col[r, :hyp_len[r] + 1], wcol)


def test_wavefront_large_tokens(cuda):
    from pgasr_b200 import functional as F
    rng = np.random.default_rng(5)
    refs = rng.integers(-2**31, 2**31 - 1, (3, 700), dtype=np.int64).astype(np.int32)
    hyps = refs.copy()[:, :650]
    hyps[:, ::7] += 1                                                    # substitutions
    d = F.edit_distance_tokens(dev_t(hyps, cuda), None, dev_t(refs, cuda), None).cpu().numpy()
    for i in range(3):
        assert d[i] == cport.edit_distance(refs[i], hyps[i])


# ---------------------------------------------------------------- a7 policy gradient
@pytest.mark.parametrize("reward,baseline", [("ed", "mean"), ("cer", "loo"), ("cer", "none"), ("ed", "value")])
def test_pg_advantages_and_grad(cuda, reward, baseline):
    from pgasr_b200 import functional as F
    B, T, V, K, L = 5, 61, 30, 16, 12
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=11, ragged=True)
    samples, logp = cport.softmax_sample(logits, in_len, uni)
    _, _, dist = cport.collapse_score(samples, targets, in_len, tgt_len)
    rm, bm = F.REWARD_MODES[reward], F.BASELINE_MODES[baseline]
    loss, R, A, grad = cport.pg_loss_grad(logits, samples, logp, dist, in_len, tgt_len, L, rm, bm, -3.5)
    R_g, A_g, terms = F.pg_advantages(dev_t(dist, cuda), dev_t(tgt_len, cuda), dev_t(logp.astype(np.float32), cuda),
                                      reward=reward, baseline=baseline, baseline_value=-3.5, Lmax=L)
    assert np.array_equal(R_g.cpu().numpy(), R)                           # one fp32 division: bit-exact
    assert rel_err(A_g.cpu().numpy(), A) < RTOL
    scale = np.abs(A * logp).sum() / (B * K)                            # the terms cancel; judge against their size
    assert abs(float(terms.double().sum()) / (B * K) - loss) <= RTOL * scale
    probs = torch.softmax(dev_t(logits, cuda), -1)
    g = F.pg_grad(dev_t(samples, cuda), A_g, dev_t(in_len, cuda), probs=probs if baseline != "mean" else None,
                  V=V, scale=1.0 / (B * K))
    assert rel_err(g.cpu().numpy(), grad) < RTOL


# ---------------------------------------------------------------- a8 CTC
def ctc_case(cuda, B, T, V, L, seed, ragged, regime="random", repeat=False):
    from pgasr_b200 import functional as F
    logits, targets, in_len, tgt_len, _ = make_batch(B, T, V, 1, L, seed=seed, ragged=ragged, regime=regime)
    if repeat:
        targets[0, :min(4, L)] = targets[0, 0]
    nll_ref, g_ref = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
    nll, g = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda))
    nll, g = nll.cpu().numpy(), g.cpu().numpy()
    fin = np.isfinite(nll_ref)
    assert np.array_equal(np.isfinite(nll), fin)
    assert np.abs(nll[fin] - nll_ref[fin]).max() <= RTOL * np.abs(nll_ref[fin]).max()
    assert np.abs(nll[fin] / nll_ref[fin] - 1).max() < RTOL
    assert rel_err(g, g_ref) < RTOL
    for b in range(B):
        assert (g[b, in_len[b]:] == 0).all()
    return nll, g


@pytest.mark.parametrize("B,T,V,L,ragged,regime", [(4, 40, 7, 9, False, "random"), (6, 123, 30, 20, True, "random"),
                                                   (64, 500, 30, 100, False, "random"), (8, 500, 30, 100, True, "peaky"),
                                                   (3, 64, 5, 63, True, "random"), (2, 300, 30, 127, False, "random"),
                                                   (2, 600, 30, 128, False, "random"), (2, 900, 12, 255, True, "peaky"),
                                                   (2, 1100, 30, 400, False, "random"), (3, 9, 4, 1, False, "random"),
                                                   # long utterances: tile streamed from the workspace (any T)
                                                   (3, 1203, 32, 60, True, "random"), (2, 2000, 30, 250, False, "peaky"),
                                                   (2, 1501, 2, 100, True, "random"), (4, 4000, 30, 30, True, "random")])
def test_ctc_matches_oracle(cuda, B, T, V, L, ragged, regime):
    ctc_case(cuda, B, T, V, L, seed=T + L, ragged=ragged, regime=regime, repeat=True)


def test_ctc_config3_full_size(cuda):
    """BASELINE.json configs[2]: B=128, T=1000, V=30, label length 200."""
    ctc_case(cuda, 128, 1000, 30, 200, seed=1, ragged=False)


def test_ctc_against_torch_cpu_double(cuda):
    """Independent external check: torch.nn.functional.ctc_loss on the CPU in fp64."""
    from pgasr_b200 import functional as F
    B, T, V, L = 5, 150, 30, 25
    logits, targets, in_len, tgt_len, _ = make_batch(B, T, V, 1, L, seed=77, ragged=True)
    x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    want = torch.nn.functional.ctc_loss(torch.log_softmax(x, -1).transpose(0, 1), torch.tensor(targets, dtype=torch.long),
                                        torch.tensor(in_len, dtype=torch.long), torch.tensor(tgt_len, dtype=torch.long),
                                        blank=0, reduction="none")
    want.sum().backward()
    nll, g = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda))
    assert np.abs(nll.cpu().numpy() / want.detach().numpy() - 1).max() < RTOL
    assert rel_err(g.cpu().numpy(), x.grad.numpy()) < RTOL


def test_ctc_edge_cases(cuda):
    from pgasr_b200 import functional as F
    logits, targets, in_len, tgt_len, _ = make_batch(4, 12, 6, 1, 5, seed=3)
    targets[0] = [1, 1, 1, 1, 1]; in_len[0] = 6                           # infeasible: needs 9 frames
    targets[1] = [1, 2, 3, 4, 1]; in_len[1] = 5                           # exactly feasible, single path family
    tgt_len[2] = 0                                                        # empty transcript: all-blank path
    in_len[3] = 1; tgt_len[3] = 1
    nll_ref, g_ref = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
    nll, g = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda))
    nll, g = nll.cpu().numpy(), g.cpu().numpy()
    assert math.isinf(nll[0]) and math.isinf(nll_ref[0]) and (g[0] == 0).all()
    assert np.abs(nll[1:] / nll_ref[1:] - 1).max() < RTOL
    assert rel_err(g, g_ref) < RTOL
    # accumulate into an existing gradient and scale
    base = torch.full((4, 12, 6), 0.25, device=cuda)
    _, g2 = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                            grad_scale=0.5, out=base)
    assert np.abs(g2.cpu().numpy() - (0.25 + 0.5 * g)).max() < 1e-6


# ---------------------------------------------------------------- a5 customNLLLoss
def test_custom_nll_loss(cuda, golden):
    import pgasr_b200
    g = golden["nll"]
    inp = torch.tensor(g["inp"], dtype=torch.float32, device=cuda, requires_grad=True)
    tgt = torch.tensor(g["target"], dtype=torch.long, device=cuda)
    for c in g["cases"]:
        inp.grad = None
        crit = pgasr_b200.loss.customNLLLoss(ignore_index=c["ignore_index"])
        loss = crit(inp, tgt)
        assert abs(float(loss) - c["out"]) < 1e-5
        (2.0 * loss).backward()
        ign = c["ignore_index"] if c["ignore_index"] else -1
        _, gref = cport.nll_sum(np.array(g["inp"], np.float32), np.array(g["target"]), ign, want_grad=True)
        assert np.abs(inp.grad.cpu().numpy() - 2.0 * gref).max() < 1e-6


# ---------------------------------------------------------------- whole step
def step_case(cuda, B, T, V, K, L, seed, ragged, regime, reward="ed", baseline="mean", w_pg=1.0, w_ctc=1.0, philox=False):
    from pgasr_b200 import functional as F
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=seed, ragged=ragged, regime=regime)
    kw = dict(reward_mode=F.REWARD_MODES[reward], baseline_mode=F.BASELINE_MODES[baseline], baseline_value=-2.0,
              w_pg=w_pg, w_ctc=w_ctc)
    loss_ref, R_ref, nll_ref, dl_ref = cport.pg_ctc_step(logits, targets, in_len, tgt_len, None if philox else uni,
                                                         seed=99, K=K, **kw)
    s_ref, _ = cport.softmax_sample(logits, in_len, None if philox else uni, seed=99, K=K)
    h_ref, hl_ref, d_ref = cport.collapse_score(s_ref, targets, in_len, tgt_len)
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda), K=K,
                        reward=reward, baseline=baseline, baseline_value=-2.0, pg_weight=w_pg, ctc_weight=w_ctc,
                        uniforms=None if philox else dev_t(uni, cuda), seed=99,
                        want=("rewards", "nll", "logp", "dist", "hyp_len", "samples"))
    if w_pg:
        assert np.array_equal(out["samples"].cpu().numpy(), s_ref)        # bit-exact
        assert np.array_equal(out["hyp_len"].cpu().numpy(), hl_ref)
        assert np.array_equal(out["dist"].cpu().numpy(), d_ref)
        assert np.array_equal(out["rewards"].cpu().numpy(), R_ref)
    if w_ctc:
        assert np.abs(out["nll"].cpu().numpy() / nll_ref - 1).max() < RTOL
    assert abs(float(out["loss"]) - loss_ref) <= RTOL * abs(loss_ref) + 1e-5
    assert rel_err(out["dlogits"].cpu().numpy(), dl_ref) < RTOL
    assert grad_close(out["dlogits"].cpu().numpy(), dl_ref)
    return out


@pytest.mark.parametrize("B,T,V,K,L,ragged,regime", [(3, 50, 30, 4, 8, True, "random"), (64, 500, 30, 16, 100, False, "random"),
                                                     (16, 500, 30, 16, 100, True, "peaky"), (2, 250, 30, 64, 40, False, "random"),
                                                     (2, 1000, 30, 4, 200, True, "peaky")])
def test_step_matches_oracle(cuda, B, T, V, K, L, ragged, regime):
    step_case(cuda, B, T, V, K, L, seed=B * 7 + K, ragged=ragged, regime=regime)


def test_step_modes(cuda):
    step_case(cuda, 4, 80, 30, 8, 12, 1, True, "random", reward="cer", baseline="loo", w_pg=0.3, w_ctc=0.7)
    step_case(cuda, 4, 80, 30, 8, 12, 2, True, "peaky", reward="cer", baseline="value", w_pg=1.0, w_ctc=0.0)
    step_case(cuda, 4, 80, 30, 8, 12, 3, False, "random", reward="ed", baseline="none", w_pg=0.0, w_ctc=1.0)
    step_case(cuda, 4, 80, 30, 16, 12, 4, True, "random", philox=True)


def test_step_all_blank_hypotheses(cuda):
    """Every sample collapses to the empty string: reward = -len(ref), advantage 0."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 2, 40, 30, 4, 6
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=1)
    logits[:, :, 0] += 80.0
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                        uniforms=dev_t(uni, cuda), want=("rewards", "hyp_len", "dist"))
    assert (out["hyp_len"].cpu().numpy() == 0).all()
    assert (out["dist"].cpu().numpy() == L).all()
    assert (out["rewards"].cpu().numpy() == -float(L)).all()


def test_module_autograd_slot(cuda):
    """PolicyGradCTCLoss in the upstream criterion slot: loss = criterion(model_out, t); loss.backward()."""
    import pgasr_b200
    B, T, V, K, L = 4, 60, 30, 8, 10
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=8, ragged=True)
    feats = torch.randn(B, T, 16, device=cuda)
    head = torch.nn.Linear(16, V).to(cuda)
    model_out = head(feats)
    crit = pgasr_b200.PolicyGradCTCLoss(K=K, reward="cer", pg_weight=0.5, ctc_weight=1.0)
    loss = crit(model_out, dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda), uniforms=dev_t(uni, cuda))
    loss.backward()
    loss_ref, _, _, dl_ref = cport.pg_ctc_step(model_out.detach().cpu().numpy(), targets, in_len, tgt_len, uni,
                                               reward_mode=1, w_pg=0.5, w_ctc=1.0)
    assert abs(float(loss) - loss_ref) <= RTOL * abs(loss_ref)
    want_w = torch.einsum("btv,btf->vf", torch.tensor(dl_ref), feats.cpu())
    assert rel_err(head.weight.grad.cpu().numpy(), want_w.numpy()) < 5e-4
    assert crit.last["rewards"].shape == (B, K)
    with pytest.raises(ZeroDivisionError):
        crit(model_out, dev_t(targets, cuda), dev_t(in_len, cuda), torch.zeros(B, dtype=torch.int32))


def test_size_independent_properties_full_size(cuda):
    """BASELINE.json configs[1] size: properties that hold without the oracle."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 64, 500, 30, 16, 100
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=0)
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), None, None, uniforms=dev_t(uni, cuda),
                        want=("rewards", "hyp_len", "dist", "samples", "nll"))
    d, hl = out["dist"].cpu().numpy(), out["hyp_len"].cpu().numpy()
    assert (d >= np.abs(hl - L)).all() and (d <= np.maximum(hl, L)).all()      # Levenshtein bounds
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    assert np.abs(g.sum(-1)).max() < 1e-5            # both gradients are differences of distributions: rows sum to 0
    # the repeat merge (upstream collapse_fn) is idempotent; the blank drop after it only shortens
    s = out["samples"]
    m1, k1 = F.collapse(s, blank=None)
    m2, k2 = F.collapse(m1, k1.reshape(-1), blank=None)
    assert torch.equal(k1, k2) and torch.equal(m1, m2)
    c1, n1 = F.collapse(s, blank=0)
    assert bool((n1 <= k1).all()) and torch.equal(n1, out["hyp_len"])
    same = F.edit_distance(c1[:, 0, :].contiguous(), n1[:, 0].contiguous(), c1[:, 0, :].to(torch.int32).contiguous(),
                           n1[:, 0].contiguous(), rows_per_ref=1, vocab=V)
    assert (same.cpu().numpy() == 0).all()
    # determinism: the same call twice is bit-identical
    out2 = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), None, None, uniforms=dev_t(uni, cuda), want=("nll",))
    assert torch.equal(out["dlogits"], out2["dlogits"]) and torch.equal(out["nll"], out2["nll"])


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [1, 3])
def test_host_pipeline_matches_device_step(cuda, depth):
    """pgasr_host_* on pinned HOST arrays: every step's outputs equal the device-tensor step bit for bit (same
    Philox seed), with several steps in flight, slots reused, ragged lengths and the oracle as the checker."""
    import pgasr_b200
    from pgasr_b200 import functional as F
    B, T, V, K, L = 6, 90, 30, 8, 14
    nsteps = 2 * depth + 3
    batches = [make_batch(B, T, V, K, L, seed=100 + i, ragged=(i % 2 == 0)) for i in range(nsteps)]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    with pgasr_b200.HostPipeline(B, T, V, K, L, depth=depth, reward="cer", baseline="loo", pg_weight=0.7,
                                 ctc_weight=1.3) as pipe:
        outs, tickets, keep = [], [], []
        for i, (lg, tg, il, tl, _) in enumerate(batches):
            h = (pin(lg), pin(tg), pin(il), pin(tl))
            keep.append(h)
            o = pipe.output_buffers()
            tickets.append(pipe.submit(*h, out=o, seed=7 + i))
            outs.append(o)
        assert tickets == list(range(nsteps))
        pipe.wait(tickets[1])                       # partial wait, then everything
        pipe.wait()
        # full-length variant through the synchronous convenience call
        lg, tg, _, _, _ = batches[0]
        o_full = pipe.step(pin(lg), pin(tg), seed=3)
    for i, (lg, tg, il, tl, _) in enumerate(batches):
        ref = F.pg_ctc_step(dev_t(lg, cuda), dev_t(tg, cuda), dev_t(il, cuda), dev_t(tl, cuda), K=K, reward="cer",
                            baseline="loo", pg_weight=0.7, ctc_weight=1.3, seed=7 + i, want=("rewards", "nll", "samples"))
        assert torch.equal(outs[i]["dlogits"], ref["dlogits"].cpu()), i
        assert torch.equal(outs[i]["rewards"], ref["rewards"].cpu())
        assert torch.equal(outs[i]["nll"], ref["nll"].cpu())
        assert float(outs[i]["loss"][0]) == float(ref["loss"])
        if i == 0:                                   # and the oracle agrees (Philox seed 7)
            loss_ref, R_ref, nll_ref, dl_ref = cport.pg_ctc_step(lg, tg, il, tl, None, seed=7, K=K, reward_mode=1,
                                                                 baseline_mode=2, w_pg=0.7, w_ctc=1.3)
            assert np.array_equal(outs[0]["rewards"].numpy(), R_ref)
            assert rel_err(outs[0]["dlogits"].numpy(), dl_ref) < RTOL
            assert abs(float(outs[0]["loss"][0]) - loss_ref) <= RTOL * abs(loss_ref)
    ref = F.pg_ctc_step(dev_t(batches[0][0], cuda), dev_t(batches[0][1], cuda), None, None, K=K, reward="cer",
                        baseline="loo", pg_weight=0.7, ctc_weight=1.3, seed=3)
    assert torch.equal(o_full["dlogits"], ref["dlogits"].cpu())


@pytest.mark.gpu
def test_host_pipeline_rejects_bad_arguments(cuda):
    import pgasr_b200
    from pgasr_b200 import _native
    with pytest.raises(_native.PgasrError):
        pgasr_b200.HostPipeline(4, 50, 65, 8, 10)             # V > 64: unsupported
    with pgasr_b200.HostPipeline(2, 20, 30, 4, 5, depth=2) as pipe:
        with pytest.raises(TypeError):
            pipe.submit(torch.zeros(2, 20, 30, device=cuda), torch.zeros(2, 5, dtype=torch.int32))
        with pytest.raises(ValueError):
            pipe.submit(torch.zeros(2, 19, 30), torch.zeros(2, 5, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        pipe.submit(torch.zeros(2, 20, 30), torch.zeros(2, 5, dtype=torch.int32))


# ------------------------------------------------------------------------------- prefix beam search
def _softmax_rows(z):
    p = np.exp(z - z.max(-1, keepdims=True))
    return p / p.sum(-1, keepdims=True)


@pytest.mark.gpu
def test_beam_search_upstream_golden(cuda, golden):
    """CTCDecoder.decode against the vectors produced by the real upstream decoder (CTCdecoder.py:41-116)."""
    import pgasr_b200
    dec = pgasr_b200.CTCdecoder.CTCDecoder(alphabet=None)
    for e in golden["beam_search"]:
        labels, nll = dec.decode(np.array(e["probs"]), beam_size=e["beam"])
        assert list(labels) == e["labels"]
        assert abs(nll - e["nll"]) < 1e-9
    # several utterances (of different lengths) in one launch
    e = golden["beam_search"][2]
    full = np.array(e["probs"])
    outs = dec.decode_batch([full, full[:7], full], beam_size=e["beam"])
    assert list(outs[0][0]) == e["labels"] and list(outs[2][0]) == e["labels"] and abs(outs[2][1] - e["nll"]) < 1e-9
    assert outs[1] == dec.decode(full[:7], beam_size=e["beam"])


@pytest.mark.gpu
@pytest.mark.parametrize("T,V,beam,peaky", [(40, 5, 1, False), (60, 30, 5, False), (80, 30, 5, True), (50, 8, 100, False),
                                             (35, 30, 128, True), (120, 4, 16, False), (30, 64, 7, False)])
def test_beam_search_matches_oracle_random(cuda, T, V, beam, peaky):
    """Random posteriors (incl. exact zeros and ragged lengths) against the oracle restatement of upstream's decoder."""
    from oracle import pyref
    from pgasr_b200 import functional as F
    rng = np.random.default_rng(T * 1000 + V * 10 + beam)
    N = 5
    z = rng.normal(size=(N, T, V)) * (4.0 if peaky else 1.5)
    if peaky:
        z[:, :, 0] += 3.0
    p = _softmax_rows(z)
    p[0, ::7, 1] = 0.0                                      # log(0) = -inf entries (rows need not sum to one)
    lens = np.array([T, T - 1, max(T // 2, 1), 1, T], np.int32)
    labels, label_len, nll = F.ctc_beam_search(dev_t(p, cuda), dev_t(lens, cuda), beam_size=beam)
    labels, label_len, nll = labels.cpu().numpy(), label_len.cpu().numpy(), nll.cpu().numpy()
    for n in range(N):
        with np.errstate(divide="ignore"):
            ref_labels, ref_nll = pyref.prefix_beam_search(p[n, :lens[n]], beam_size=beam)
        assert tuple(labels[n, :label_len[n]]) == tuple(ref_labels), n
        assert abs(nll[n] - ref_nll) <= 1e-9 * max(1.0, abs(ref_nll)), n
        assert (labels[n, label_len[n]:] == 0).all()


@pytest.mark.gpu
def test_beam_search_properties_full_size(cuda):
    """T=500, V=30 (upstream's reward() and predict() call it with beam_size=5): beam=1 on one-hot posteriors is
    the collapsed argmax path; the beam score is bounded by the exact CTC likelihood of the returned labels."""
    from pgasr_b200 import functional as F
    rng = np.random.default_rng(3)
    N, T, V = 16, 500, 30
    path = rng.integers(0, V, size=(N, T))
    onehot = np.full((N, T, V), 1e-12)
    np.put_along_axis(onehot, path[..., None], 1.0, axis=-1)
    labels, label_len, _ = F.ctc_beam_search(dev_t(onehot, cuda), None, beam_size=1)
    col, col_len = F.collapse(dev_t(path.astype(np.uint8), cuda), blank=0)
    for n in range(N):
        k = int(label_len[n])
        assert k == int(col_len[n]) and torch.equal(labels[n, :k].to(torch.uint8), col[n, :k])
    # the beam sums a subset of the alignments of its best prefix: its score can never beat the exact CTC
    # likelihood of that label sequence (computed by the CTC kernel of this library)
    p = _softmax_rows(rng.normal(size=(N, T, V)) * 2)
    logits = dev_t(np.log(p).astype(np.float32), cuda)
    for beam in (1, 5, 25):
        labels, label_len, nll = F.ctc_beam_search(dev_t(p, cuda), None, beam_size=beam)
        assert bool(torch.isfinite(nll).all()) and int(label_len.min()) > 0
        Lmax = int(label_len.max())
        exact, _ = F.ctc_loss_grad(logits, labels[:, :Lmax].contiguous(), None, label_len)
        assert bool((nll.cpu() >= exact.cpu().double() * (1 - 1e-4)).all())
    with pytest.raises(Exception):
        F.ctc_beam_search(dev_t(p, cuda), None, beam_size=129)


@pytest.mark.gpu
def test_evaluate_batch_matches_upstream_golden(cuda, golden):
    """the predict() loop's per-utterance evaluate() (metrics.py:23-31) as one batched call."""
    import pgasr_b200
    ev = golden["evaluate"]
    got = pgasr_b200.metrics.evaluate_batch([e["ref"] for e in ev], [e["hyp"] for e in ev])
    for g, e in zip(got, ev):
        assert g == tuple(e["out"])                  # int / int in Python floats on both sides: exact
        assert pgasr_b200.metrics.evaluate(e["ref"], e["hyp"]) == tuple(e["out"])
    with pytest.raises(ZeroDivisionError):
        pgasr_b200.metrics.evaluate_batch(["ab", ""], ["ab", "x"])


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,K,L", [(64, 500, 30, 16, 100), (24, 1000, 30, 8, 200), (148, 96, 30, 4, 20)])
def test_step_soak_bit_reproducible(cuda, B, T, V, K, L):
    """The walker / worker hand-off (named barriers, rings, flags between CTAs) under repetition: the same step
    300 times must give bit-identical gradients every time (a lost or early hand-off would show as a diff)."""
    from pgasr_b200 import functional as F
    lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=11, ragged=True)
    lg, tg, il, tl = dev_t(lg, cuda), dev_t(tg, cuda), dev_t(il, cuda), dev_t(tl, cuda)
    first = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=5, want=("nll", "rewards"))
    ws = first["workspace"]
    bad = torch.zeros((), dtype=torch.int64, device=cuda)
    for i in range(300):
        out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=5, workspace=ws, want=("nll", "rewards"))
        bad += (out["dlogits"] != first["dlogits"]).sum() + (out["nll"] != first["nll"]).sum() + \
            (out["rewards"] != first["rewards"]).sum() + (out["loss"] != first["loss"]).sum()
    assert int(bad) == 0


@pytest.mark.gpu
def test_predict_inner_loop_matches_upstream_semantics(cuda):
    """predict.decode_and_score against upstream's per-utterance loop (model.py:321-334) replayed with the oracle."""
    import pgasr_b200
    from oracle import pyref
    rng = np.random.default_rng(5)
    B, T, V, L = 6, 40, 8, 12
    ind2char = {0: "_", 1: " ", 2: "a", 3: "b", 4: "c", 5: "d", 6: "e", 7: "l"}
    z = rng.normal(size=(B, T, V)) * 2.5
    logp = z - np.log(np.exp(z).sum(-1, keepdims=True))
    t = rng.integers(1, V, size=(B, L))
    t[:, 0] = 2                                                  # no leading space: every reference has a first word
    tmask = np.zeros((B, L), np.int64)
    for i in range(B):
        tmask[i, :rng.integers(3, L + 1)] = 1
    targets, predicted, cers, wers = pgasr_b200.predict.decode_and_score(logp, t, tmask, ind2char, beam_size=5)
    for i in range(B):
        pad = int(tmask[i].sum())
        with np.errstate(divide="ignore"):
            seq, _ = pyref.prefix_beam_search(np.exp(logp[i][:pad]), beam_size=5)
        hyp = pyref.collapse_fn("".join(ind2char[s] for s in seq))
        tgt = "".join(ind2char[s] for s in t[i][:pad])
        assert targets[i] == tgt and predicted[i] == hyp
        cer, wer = pyref.evaluate(tgt, hyp)
        assert cers[i] == cer and wers[i] == wer


@pytest.mark.gpu
def test_acoustic_harness_trains(cuda):
    """the criterion slot in a real training step (stock BLSTM + head): finite loss, parameters move."""
    import subprocess, sys, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "acoustic_harness.py"), "--global-batch", "8", "--T", "120",
                        "--L", "20", "--K", "4", "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert np.isfinite(out["loss"]) and out["ms_per_step"] > 0 and -1.5 * 120 / 20 <= out["mean_reward"] <= 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,K,L,philox", [(1, 33, 30, 1, 4, False), (2, 70, 9, 5, 11, True), (3, 65, 32, 7, 6, False),
                                               (2, 500, 30, 64, 100, True), (5, 48, 3, 2, 47, False),
                                               (3, 120, 33, 5, 20, False), (2, 500, 40, 16, 100, True), (2, 260, 64, 8, 70, False),
                                               (1, 1300, 64, 4, 150, False), (2, 90, 47, 3, 200, True)])
def test_step_odd_shapes(cuda, B, T, V, K, L, philox):
    """K not a multiple of the Philox block, single utterance / single sample, V = 32 (row stride 34), tiny V, and
    alphabets wider than 32 classes (33 / 40 / 47 / 64: upstream's alphabet is a character set plus punctuation,
    model.py:195): samples bit-exact, gradients within 1e-4, also in the streaming modes (T = 1300) and with 16 CTC
    states per lane (L = 150 / 200)."""
    step_case(cuda, B, T, V, K, L, seed=B + K, ragged=True, regime="random", philox=philox)


@pytest.mark.gpu
def test_step_empty_utterances_and_nonzero_blank(cuda):
    """in_len = 0 (nothing to sample or align), an empty transcript, and a blank id other than 0."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 4, 40, 6, 4, 5
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=21)
    in_len[0] = 0                                   # no frames: every hypothesis is empty, CTC has no alignment
    tgt_len[1] = 0; targets[1] = 0                  # empty transcript: only the all-blank path
    in_len[2] = 3; tgt_len[2] = 5                   # fewer frames than labels: infeasible
    loss_ref, R_ref, nll_ref, dl_ref = cport.pg_ctc_step(logits, targets, in_len, tgt_len, uni, w_pg=1.0, w_ctc=0.0)
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                        uniforms=dev_t(uni, cuda), ctc_weight=0.0, want=("rewards", "hyp_len", "dist"))
    assert np.array_equal(out["rewards"].cpu().numpy(), R_ref)
    assert rel_err(out["dlogits"].cpu().numpy(), dl_ref) < RTOL       # (the oracle's scalar is 0 * inf = nan here)
    assert np.isfinite(float(out["loss"]))
    assert (out["hyp_len"][0].cpu().numpy() == 0).all() and (out["dist"][0].cpu().numpy() == L).all()
    nll_ref, g_ref = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                        uniforms=dev_t(uni, cuda), pg_weight=0.0, want=("nll",))
    nll = out["nll"].cpu().numpy()
    assert np.isinf(nll[0]) and np.isinf(nll[2]) and np.array_equal(np.isfinite(nll), np.isfinite(nll_ref))
    fin = np.isfinite(nll_ref)
    assert np.abs(nll[fin] / nll_ref[fin] - 1).max() < RTOL
    g = out["dlogits"].cpu().numpy() * B
    assert (g[0] == 0).all() and (g[2] == 0).all() and rel_err(g[fin], g_ref[fin]) < RTOL
    # blank id 2: labels avoid it
    logits, targets, in_len, tgt_len, uni = make_batch(3, 50, 6, 4, 7, seed=22, ragged=True)
    targets = np.where(targets == 2, 5, targets).astype(np.int32)
    for b in range(3):
        targets[b, tgt_len[b]:] = 0
    loss_ref, R_ref, nll_ref, dl_ref = cport.pg_ctc_step(logits, targets, in_len, tgt_len, uni, blank=2)
    out = F.pg_ctc_step(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                        uniforms=dev_t(uni, cuda), blank=2, want=("rewards", "nll"))
    assert np.array_equal(out["rewards"].cpu().numpy(), R_ref)
    assert np.abs(out["nll"].cpu().numpy() / nll_ref - 1).max() < RTOL
    assert rel_err(out["dlogits"].cpu().numpy(), dl_ref) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,K,L,reward,baseline", [(2, 1000, 30, 64, 120, "ed", "mean"), (2, 2000, 30, 16, 100, "cer", "loo"),
                                                        (3, 1300, 12, 8, 250, "ed", "none"), (2, 2000, 30, 4, 400, "ed", "mean")])
def test_step_long_utterances_single_launch(cuda, B, T, V, K, L, reward, baseline):
    """Long T / large K: both roles stream (no [T][V] tile in shared memory), still one launch, same parity bar."""
    from pgasr_b200 import _native
    n0 = _native.lib().pgasr_launch_count()
    step_case(cuda, B, T, V, K, L, seed=T + K, ragged=True, regime="random", reward=reward, baseline=baseline)
    assert _native.lib().pgasr_launch_count() - n0 == 1


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,K,L", [(64, 500, 30, 16, 100), (8, 1000, 30, 4, 200), (3, 77, 9, 3, 12)])
def test_step_does_not_depend_on_stale_workspace(cuda, B, T, V, K, L):
    """Every lattice row a gradient worker reads must have been written in THIS launch: the workspace is poisoned
    (all bits set = NaN) between steps, so a row read before the other direction stored it would surface as NaN or
    as a difference from the clean run (the soak test alone cannot see it: identical steps leave identical rows)."""
    from pgasr_b200 import functional as F
    lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=31, ragged=True)
    lg, tg, il, tl = dev_t(lg, cuda), dev_t(tg, cuda), dev_t(il, cuda), dev_t(tl, cuda)
    ref = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=5, want=("nll",))
    ws = ref["workspace"]
    for i in range(20):
        ws.buf[65536:].fill_(0xFF)                       # the control block at the front stays armed
        out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=5, workspace=ws, want=("nll",))
        assert bool(torch.isfinite(out["dlogits"]).all())
        assert torch.equal(out["dlogits"], ref["dlogits"]) and torch.equal(out["nll"], ref["nll"])


@pytest.mark.gpu
def test_step_randomised_shapes_and_modes(cuda):
    """tools/fuzz_step.py: random (B, T, V, K, L, reward, baseline, weights, regime, Philox or injected uniforms) across
    the boundaries between the kernel's modes, every case against the C oracle with the usual bars."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_step.py"), "200", "5"], capture_output=True, text=True,
                       timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all 200 cases passed" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V,L", [(4, 40, 7, 9), (2, 900, 12, 255), (2, 1100, 30, 400), (3, 2500, 30, 20), (2, 300, 40, 50)])
def test_ctc_classic_kernel_path(cuda, B, T, V, L):
    """The one-CTA-per-utterance kernel (taken for `accumulate`, for V > 32 and for probabilities-only input) keeps the
    same bar as the walker/worker kernel, including T >> L where alpha * beta underflows fp64 unless the scaling is split."""
    from pgasr_b200 import functional as F
    logits, targets, in_len, tgt_len, _ = make_batch(B, T, V, 1, L, seed=T + V, ragged=True)
    nll_ref, g_ref = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
    base = torch.full((B, T, V), 0.25, dtype=torch.float32, device=cuda)
    nll, g = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
                             grad_scale=0.5, out=base)
    assert g.data_ptr() == base.data_ptr()
    nll, g = nll.cpu().numpy(), (g.cpu().numpy() - 0.25) / 0.5
    fin = np.isfinite(nll_ref)
    assert np.array_equal(np.isfinite(nll), fin) and np.abs(nll[fin] / nll_ref[fin] - 1).max() < RTOL
    assert np.abs(g - g_ref).max() / np.abs(g_ref).max() < 2 * RTOL      # (the 0.25 offset costs a few fp32 ulps)


@pytest.mark.gpu
def test_step_hybrid_path_when_samples_do_not_fit(cuda):
    """K = 64 at T = 1800: the sample buffers (2 K T bytes) exceed one SM, so the PG part runs as the stand-alone
    kernels next to the single-launch CTC role; same parity bar, more than one launch."""
    from pgasr_b200 import _native
    n0 = _native.lib().pgasr_launch_count()
    step_case(cuda, 2, 1800, 30, 64, 50, seed=9, ragged=True, regime="random", reward="cer", baseline="loo")
    assert _native.lib().pgasr_launch_count() - n0 > 1


@pytest.mark.gpu
def test_step_back_to_back_distinct_batches(cuda):
    """Programmatic dependent launch: steps on DIFFERENT batches enqueued back to back on one workspace (no host sync in
    between, so the CTAs of step n+1 are placed while step n runs) must each give exactly what the same step gives
    alone on a fresh workspace."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 64, 500, 30, 16, 100
    batches = []
    for s in range(5):
        lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=100 + s, ragged=(s % 2 == 1))
        batches.append(tuple(dev_t(x, cuda) for x in (lg, tg, il, tl)))
    refs = []
    for s, (lg, tg, il, tl) in enumerate(batches):
        out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=s, want=("rewards", "nll"))
        torch.cuda.synchronize()
        refs.append({k: out[k].clone() for k in ("loss", "dlogits", "rewards", "nll")})
    ws = None
    outs = []
    for rep in range(6):
        for s, (lg, tg, il, tl) in enumerate(batches):
            out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=s, workspace=ws, want=("rewards", "nll"))
            ws = out["workspace"]
            outs.append((s, out))
    torch.cuda.synchronize()
    for s, out in outs:
        for k in ("loss", "dlogits", "rewards", "nll"):
            assert torch.equal(out[k], refs[s][k]), (s, k)


# ---------------------------------------------------------------- round 2
@pytest.mark.gpu
def test_step_stress_corner_all_three(cuda):
    """K = 64 AND T = 2000 AND L = 400 together (BASELINE.json configs[4]'s far corner), B = 8."""
    step_case(cuda, 8, 2000, 30, 64, 400, seed=64, ragged=True, regime="random", reward="cer", baseline="loo")


@pytest.mark.gpu
def test_step_queue_equals_single_steps(cuda):
    """pgasr_pg_ctc_step_multi (functional.StepQueue): n steps enqueued by one C-ABI call give bit for bit what the
    same steps give one call at a time, in any window of the cycle, and allocate nothing."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 16, 200, 30, 8, 40
    batches = []
    for s in range(5):
        lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=300 + s, ragged=(s % 2 == 0))
        batches.append({"logits": dev_t(lg, cuda), "targets": dev_t(tg, cuda), "in_len": dev_t(il, cuda),
                        "tgt_len": dev_t(tl, cuda)})
    q = F.StepQueue(batches, K=K, reward="cer", baseline="loo", want=("rewards", "nll", "dist"))
    for first, n, seed in [(0, 5, 7), (3, 4, 1000), (4, 1, 5), (2, 0, 0)]:
        refs = []
        for j in range(n):
            b = batches[(first + j) % 5]
            o = F.pg_ctc_step(b["logits"], b["targets"], b["in_len"], b["tgt_len"], K=K, reward="cer", baseline="loo",
                              seed=seed + j, want=("rewards", "nll", "dist"))
            refs.append({k: o[k].clone() for k in ("loss", "dlogits", "rewards", "nll", "dist")})
        torch.cuda.synchronize()
        mem0 = torch.cuda.memory_allocated()
        assert q.run(first=first, n=n, seed=seed) == n
        assert torch.cuda.memory_allocated() == mem0
        torch.cuda.synchronize()
        for j in range(n):
            o = q.outputs[(first + j) % 5]
            for k in ("dlogits", "rewards", "nll", "dist"):
                assert torch.equal(o[k], refs[j][k]), (first, j, k)
            assert torch.equal(o["loss"][0], refs[j]["loss"])
    with pytest.raises(ValueError):
        q.run(n=6)


@pytest.mark.gpu
def test_step_argument_validation(cuda):
    from pgasr_b200 import functional as F
    B, T, V, K, L = 2, 30, 6, 3, 4
    lg, tg, il, tl, uni = make_batch(B, T, V, K, L, seed=3)
    lg, tg, il, tl = dev_t(lg, cuda), dev_t(tg, cuda), dev_t(il, cuda), dev_t(tl, cuda)
    with pytest.raises(ValueError):                      # uniforms of another batch size / length: no out-of-bounds read
        F.pg_ctc_step(lg, tg, il, tl, uniforms=dev_t(uni[:1], cuda))
    with pytest.raises(ValueError):
        F.pg_ctc_step(lg, tg, il, tl, uniforms=dev_t(uni[:, :, :T - 1].copy(), cuda))
    with pytest.raises(ValueError):                      # an unknown output name is not an uninitialised tensor
        F.pg_ctc_step(lg, tg, il, tl, K=K, want=("rewards", "advantages"))
    with pytest.raises(ValueError):
        F.pg_ctc_step(lg, tg, il, tl, K=65)
    out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=1, want=("nll",))
    ws = out["workspace"]
    again = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=1, workspace=ws, out=ws.outputs(("nll",)))
    assert torch.equal(again["dlogits"], out["dlogits"])


@pytest.mark.gpu
def test_workspace_init_resets_the_parity(cuda):
    """pgasr_pg_ctc_step_workspace_init drops the host-side record of which control block comes next: a workspace
    re-initialised after an odd number of steps (or a freed pointer handed out again) behaves like a fresh one."""
    from pgasr_b200 import functional as F, _native
    B, T, V, K, L = 8, 100, 30, 4, 20
    lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=77)
    lg, tg, il, tl = dev_t(lg, cuda), dev_t(tg, cuda), dev_t(il, cuda), dev_t(tl, cuda)
    ref = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=2, want=("nll",))
    ws = ref["workspace"]
    for n_before in (1, 2, 3):
        for _ in range(n_before):
            F.pg_ctc_step(lg, tg, il, tl, K=K, seed=2, workspace=ws, want=("nll",))
        torch.cuda.synchronize()
        ws.buf.fill_(0xFF)                               # as if the block had been freed and reused by someone else
        _native.call("pgasr_pg_ctc_step_workspace_init", ws.buf.data_ptr(), ws.nbytes, torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            out = F.pg_ctc_step(lg, tg, il, tl, K=K, seed=2, workspace=ws, want=("nll",))
            assert torch.equal(out["dlogits"], ref["dlogits"]) and torch.equal(out["nll"], ref["nll"])


@pytest.mark.gpu
def test_module_takes_lengths_from_the_padding(cuda):
    """criterion(model_out, t) with zero-padded transcripts and NO lengths (the upstream slot, model.py:235): the pad
    is not a label.  Must equal the call with explicit lengths; the Lmax fallback would differ."""
    import pgasr_b200
    B, T, V, K, L = 4, 80, 30, 8, 12
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=19, ragged=True)
    assert (tgt_len < L).any()
    lg = dev_t(logits, cuda)
    crit = pgasr_b200.PolicyGradCTCLoss(K=K, reward="cer")
    a = crit(lg.clone().requires_grad_(True), dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda),
             uniforms=dev_t(uni, cuda))
    ra = crit.last["rewards"].clone()
    b = crit(lg.clone().requires_grad_(True), dev_t(targets, cuda).long(), dev_t(in_len, cuda), None, uniforms=dev_t(uni, cuda))
    assert torch.equal(a.detach(), b.detach()) and torch.equal(ra, crit.last["rewards"])
    loss_ref, R_ref, _, _ = cport.pg_ctc_step(logits, targets, in_len, tgt_len, uni, reward_mode=1)
    assert np.array_equal(ra.cpu().numpy(), R_ref) and abs(float(a) - loss_ref) <= RTOL * abs(loss_ref)


@pytest.mark.gpu
def test_out_of_range_labels_and_targets_are_loud(cuda):
    """A label id >= V (or negative) inside the transcript: no read of the neighbouring row -- the utterance has no
    alignment (nll = +inf, zero gradient), the others are untouched.  customNLLLoss: torch's ignore_index = -100 is
    honoured, a class id outside [0,V) turns the loss NaN instead of reading out of bounds."""
    import pgasr_b200
    from pgasr_b200 import functional as F
    B, T, V, L = 3, 60, 9, 7
    logits, targets, in_len, tgt_len, _ = make_batch(B, T, V, 1, L, seed=2)
    nll_ref, g_ref = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
    bad = targets.copy()
    bad[1, 3] = V + 5
    nll, g = F.ctc_loss_grad(dev_t(logits, cuda), dev_t(bad, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda))
    nll, g = nll.cpu().numpy(), g.cpu().numpy()
    assert np.isinf(nll[1]) and (g[1] == 0).all()
    assert np.abs(nll[[0, 2]] / nll_ref[[0, 2]] - 1).max() < RTOL and rel_err(g[[0, 2]], g_ref[[0, 2]]) < RTOL
    inp = torch.log_softmax(torch.randn(5, 4, 6, device=cuda), -1)
    tgt = torch.randint(0, 6, (4, 5), device=cuda)
    tgt[1, 2] = -100
    want = sum(torch.nn.functional.nll_loss(inp[i], tgt[:, i], ignore_index=-100) for i in range(5))
    got = pgasr_b200.loss.customNLLLoss(ignore_index=-100)(inp, tgt)
    assert abs(float(got) - float(want)) < 1e-5
    tgt[0, 0] = 6
    assert math.isnan(float(pgasr_b200.loss.customNLLLoss()(inp, tgt)))


# ---------------------------------------------------------------- 8f.1 reward-to-go in the training step
@pytest.mark.parametrize("B,T,V,K,L,baseline,ragged,wc", [
    (3, 120, 30, 8, 20, "mean", True, 0.0), (2, 500, 30, 16, 100, "loo", False, 1.0), (4, 77, 5, 5, 9, "none", True, 0.7),
    (2, 300, 17, 33, 40, "value", True, 0.0), (1, 64, 30, 4, 64, "mean", False, 1.0), (64, 500, 30, 16, 100, "mean", False, 1.0)])
def test_step_reward_to_go(cuda, B, T, V, K, L, baseline, ragged, wc):
    """reward='ed_to_go' (SURVEY 8f.1): r_pos and to_go bit-exact against the oracle (whose r_pos is pinned to upstream's
    policy_grad.reward by tests/test_oracle.py), rewards exact, loss and gradient within 1e-4 -- in ONE launch."""
    from pgasr_b200 import _native, functional as F
    logits, targets, in_len, tgt_len, uni = make_batch(B, T, V, K, L, seed=B * 7 + T + K, ragged=ragged)
    d = lambda a: dev_t(a, cuda)
    kw = dict(uniforms=d(uni), reward="ed_to_go", baseline=baseline, baseline_value=1.25, pg_weight=1.0, ctc_weight=wc,
              want=("rewards", "samples", "to_go", "r_pos", "nll", "dist", "hyp_len"))
    out = F.pg_ctc_step(d(logits), d(targets), d(in_len), d(tgt_len), **kw)
    n0 = _native.lib().pgasr_launch_count()
    out = F.pg_ctc_step(d(logits), d(targets), d(in_len), d(tgt_len), workspace=out["workspace"], **kw)
    assert _native.lib().pgasr_launch_count() - n0 == 1
    samples = out["samples"].cpu().numpy()
    s_ref, _ = cport.softmax_sample(logits, in_len, uni)
    assert np.array_equal(samples, s_ref)
    bm = F.BASELINE_MODES[baseline]
    loss_pg, R, to_go, r_pos, g_pg = cport.pg_togo_loss_grad(logits, samples, targets, in_len, tgt_len,
                                                             baseline_mode=bm, baseline_value=1.25)
    assert np.array_equal(out["to_go"].cpu().numpy(), to_go)
    assert np.array_equal(out["r_pos"].cpu().numpy(), r_pos)
    assert np.array_equal(out["rewards"].cpu().numpy(), R)
    want_loss, want_g = loss_pg, g_pg
    if wc:
        nll_ref, g_ctc = cport.ctc_loss_grad(logits, targets, in_len, tgt_len)
        fin = np.isfinite(nll_ref)
        g_ctc[~fin] = 0.0
        want_g = g_pg + (wc / B) * g_ctc
        want_loss = loss_pg + wc * nll_ref.mean()
    got = out["dlogits"].cpu().numpy()
    assert rel_err(got, want_g) < RTOL
    assert grad_close(got, want_g)
    if np.isfinite(want_loss):
        # the terms of the PG loss cancel; judge the scalar against their size
        scale = max(abs(want_loss), np.abs(to_go).mean() * 4.0)
        assert abs(float(out["loss"]) - want_loss) <= RTOL * scale


def test_step_reward_to_go_through_the_loss_module(cuda):
    """PolicyGradCTCLoss(reward='ed_to_go') in the criterion(model_out, t) slot: autograd hands back the fused gradient."""
    import pgasr_b200
    B, T, V, K, L = 3, 90, 30, 8, 12
    logits, targets, in_len, tgt_len, _ = make_batch(B, T, V, K, L, seed=5, ragged=True)
    crit = pgasr_b200.PolicyGradCTCLoss(K=K, reward="ed_to_go", baseline="mean", seed=3)
    x = dev_t(logits, cuda).requires_grad_(True)
    loss = crit(x, dev_t(targets, cuda), dev_t(in_len, cuda), dev_t(tgt_len, cuda))
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0


@pytest.mark.gpu
def test_step_queue_overlap_soak_bit_reproducible(cuda):
    """Consecutive steps of a multi-step call overlap on two streams / two workspace lanes (up to three grids share the
    GPU): 40 runs of a 6-step window at the headline shape must reproduce the first run bit for bit, and that run must
    equal the steps taken one at a time."""
    from pgasr_b200 import functional as F
    B, T, V, K, L = 64, 500, 30, 16, 100
    batches = []
    for s in range(6):
        lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=900 + s, ragged=(s % 3 == 0))
        batches.append({"logits": dev_t(lg, cuda), "targets": dev_t(tg, cuda), "in_len": dev_t(il, cuda),
                        "tgt_len": dev_t(tl, cuda)})
    q = F.StepQueue(batches, K=K, want=("rewards", "nll"))
    q.run(first=0, n=6, seed=11)
    torch.cuda.synchronize()
    first = [{k: o[k].clone() for k in ("loss", "dlogits", "rewards", "nll")} for o in q.outputs]
    for j, b in enumerate(batches):
        o = F.pg_ctc_step(b["logits"], b["targets"], b["in_len"], b["tgt_len"], K=K, seed=11 + j, want=("rewards", "nll"))
        for k in ("dlogits", "rewards", "nll"):
            assert torch.equal(o[k], first[j][k]), (j, k)
    for rep in range(40):
        for o in q.outputs:
            o["dlogits"].fill_(float("nan"))
        q.run(first=0, n=6, seed=11)
        torch.cuda.synchronize()
        for j, o in enumerate(q.outputs):
            for k in ("loss", "dlogits", "rewards", "nll"):
                assert torch.equal(o[k], first[j][k]), (rep, j, k)
