"""ctypes binding of the CPU oracle (oracle/pgasr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  Nothing under policy-gradient-asr_b200/ does (tests/test_no_oracle_in_product.py
greps for it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpgasr_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, a second or two).  Returns the .so path."""
    src = os.path.join(_HERE, "pgasr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i32p, u8p, f32p, f64p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                 C.POINTER(C.c_float), C.POINTER(C.c_double))
        L.orc_edit_distance.argtypes = [i32p, C.c_int, i32p, C.c_int, i32p]
        L.orc_edit_distance.restype = C.c_int
        L.orc_collapse.argtypes = [i32p, C.c_int, C.c_int, i32p]
        L.orc_collapse.restype = C.c_int
        L.orc_reward_positions.argtypes = [i32p, C.c_int, i32p, C.c_int, C.c_int, i32p]
        L.orc_reward_positions.restype = None
        L.orc_nll_sum.argtypes = [f32p, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_nll_sum.restype = C.c_double
        L.orc_exp_spec.argtypes = [C.c_float]
        L.orc_exp_spec.restype = C.c_float
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_philox4x32_10.restype = None
        L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.orc_philox_uniform.restype = C.c_float
        L.orc_softmax_sample.argtypes = [f32p, i32p, f32p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                         C.c_int, u8p, f64p]
        L.orc_softmax_sample.restype = None
        L.orc_collapse_score.argtypes = [u8p, i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, u8p, i32p, i32p]
        L.orc_collapse_score.restype = None
        L.orc_pg_loss_grad.argtypes = [f32p, i32p, u8p, f64p, i32p, i32p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, f32p, f64p, f64p]
        L.orc_pg_loss_grad.restype = C.c_double
        L.orc_pg_togo_loss_grad.argtypes = [f32p, i32p, u8p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_double, f32p, C.POINTER(C.c_int16),
                                            C.POINTER(C.c_int8), f64p]
        L.orc_pg_togo_loss_grad.restype = C.c_double
        L.orc_ctc_loss_grad.argtypes = [f32p, i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, f64p, f64p]
        L.orc_ctc_loss_grad.restype = None
        L.orc_pg_ctc_step.argtypes = [f32p, i32p, i32p, i32p, f32p, C.c_uint64, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.c_double, f32p, f64p, f32p]
        L.orc_pg_ctc_step.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def max_threads():
    return int(lib().orc_max_threads())


def set_num_threads(n):
    """Use n OpenMP threads from now on (overrides an inherited OMP_NUM_THREADS)."""
    lib().orc_set_num_threads.argtypes = [C.c_int]
    lib().orc_set_num_threads.restype = None
    lib().orc_set_num_threads(int(n))


def edit_distance(ref, hyp, last_col=False):
    ref, hyp = _i32(ref), _i32(hyp)
    col = np.zeros(len(hyp) + 1, np.int32) if last_col else None
    d = lib().orc_edit_distance(_p(ref, C.c_int32), len(ref), _p(hyp, C.c_int32), len(hyp),
                                _p(col, C.c_int32))
    return (d, col) if last_col else d


def collapse(seq, blank=-1):
    seq = _i32(seq)
    out = np.zeros(max(len(seq), 1), np.int32)
    n = lib().orc_collapse(_p(seq, C.c_int32), len(seq), int(blank), _p(out, C.c_int32))
    return out[:n].copy()


def reward_positions(ref, hyp, tmax):
    ref, hyp = _i32(ref), _i32(hyp)
    r = np.zeros(tmax + 1, np.int32)
    lib().orc_reward_positions(_p(ref, C.c_int32), len(ref), _p(hyp, C.c_int32), len(hyp), tmax,
                               _p(r, C.c_int32))
    return r


def nll_sum(inp, target, ignore_index=-1, want_grad=False):
    inp = _f32(inp)
    target = np.ascontiguousarray(target, dtype=np.int64)
    L, B, V = inp.shape
    g = np.zeros((L, B, V), np.float64) if want_grad else None
    v = lib().orc_nll_sum(_p(inp, C.c_float), _p(target, C.c_int64), L, B, V, int(ignore_index),
                          _p(g, C.c_double))
    return (v, g) if want_grad else v


def exp_spec(x):
    L = lib()
    x = np.asarray(x, np.float32)
    return np.array([L.orc_exp_spec(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def philox_uniform(seed, b, t, k):
    return float(lib().orc_philox_uniform(int(seed), b, t, k))


def softmax_sample(logits, in_len=None, uniforms=None, seed=0, K=None):
    logits = _f32(logits)
    B, T, V = logits.shape
    uniforms = _f32(uniforms)
    if K is None:
        K = uniforms.shape[1]
    in_len = _i32(in_len)
    samples = np.zeros((B, K, T), np.uint8)
    logp = np.zeros((B, K), np.float64)
    lib().orc_softmax_sample(_p(logits, C.c_float), _p(in_len, C.c_int32), _p(uniforms, C.c_float),
                             int(seed), B, T, V, K, _p(samples, C.c_uint8), _p(logp, C.c_double))
    return samples, logp


def collapse_score(samples, targets, in_len=None, tgt_len=None, blank=0):
    samples = np.ascontiguousarray(samples, np.uint8)
    B, K, T = samples.shape
    targets = _i32(targets)
    Lmax = targets.shape[1]
    in_len, tgt_len = _i32(in_len), _i32(tgt_len)
    hyps = np.zeros((B, K, T), np.uint8)
    hyp_len = np.zeros((B, K), np.int32)
    dist = np.zeros((B, K), np.int32)
    lib().orc_collapse_score(_p(samples, C.c_uint8), _p(in_len, C.c_int32), _p(targets, C.c_int32),
                             _p(tgt_len, C.c_int32), B, T, K, Lmax, int(blank),
                             _p(hyps, C.c_uint8), _p(hyp_len, C.c_int32), _p(dist, C.c_int32))
    return hyps, hyp_len, dist


def pg_loss_grad(logits, samples, logp, dist, in_len=None, tgt_len=None, Lmax=None,
                 reward_mode=0, baseline_mode=1, baseline_value=0.0, want_grad=True):
    logits = _f32(logits)
    B, T, V = logits.shape
    samples = np.ascontiguousarray(samples, np.uint8)
    K = samples.shape[1]
    logp = np.ascontiguousarray(logp, np.float64)
    dist = _i32(dist)
    in_len, tgt_len = _i32(in_len), _i32(tgt_len)
    rewards = np.zeros((B, K), np.float32)
    adv = np.zeros((B, K), np.float64)
    grad = np.zeros((B, T, V), np.float64) if want_grad else None
    loss = lib().orc_pg_loss_grad(_p(logits, C.c_float), _p(in_len, C.c_int32),
                                  _p(samples, C.c_uint8), _p(logp, C.c_double),
                                  _p(dist, C.c_int32), _p(tgt_len, C.c_int32), B, T, V, K,
                                  int(Lmax if Lmax is not None else 0), int(reward_mode),
                                  int(baseline_mode), float(baseline_value),
                                  _p(rewards, C.c_float), _p(adv, C.c_double), _p(grad, C.c_double))
    return loss, rewards, adv, grad


def pg_togo_loss_grad(logits, samples, targets, in_len=None, tgt_len=None, blank=0, baseline_mode=1,
                      baseline_value=0.0, want_grad=True):
    """per-position rewards (policy_grad.py:10-15) credited as reward-to-go.
    -> loss, rewards [B,K] (= len(ref) - ED), to_go [B,K,T] int16, r_pos [B,K,T] int8, grad [B,T,V] fp64."""
    logits = _f32(logits)
    B, T, V = logits.shape
    samples = np.ascontiguousarray(samples, np.uint8)
    K = samples.shape[1]
    targets = _i32(targets)
    Lmax = targets.shape[1]
    in_len, tgt_len = _i32(in_len), _i32(tgt_len)
    rewards = np.zeros((B, K), np.float32)
    to_go = np.zeros((B, K, T), np.int16)
    r_pos = np.zeros((B, K, T), np.int8)
    grad = np.zeros((B, T, V), np.float64) if want_grad else None
    loss = lib().orc_pg_togo_loss_grad(_p(logits, C.c_float), _p(in_len, C.c_int32), _p(samples, C.c_uint8),
                                       _p(targets, C.c_int32), _p(tgt_len, C.c_int32), B, T, V, K, Lmax,
                                       int(blank), int(baseline_mode), float(baseline_value),
                                       _p(rewards, C.c_float), _p(to_go, C.c_int16), _p(r_pos, C.c_int8),
                                       _p(grad, C.c_double))
    return loss, rewards, to_go, r_pos, grad


def ctc_loss_grad(logits, targets, in_len=None, tgt_len=None, blank=0, want_grad=True):
    logits = _f32(logits)
    B, T, V = logits.shape
    targets = _i32(targets)
    Lmax = targets.shape[1]
    in_len, tgt_len = _i32(in_len), _i32(tgt_len)
    nll = np.zeros(B, np.float64)
    grad = np.zeros((B, T, V), np.float64) if want_grad else None
    lib().orc_ctc_loss_grad(_p(logits, C.c_float), _p(targets, C.c_int32), _p(in_len, C.c_int32),
                            _p(tgt_len, C.c_int32), B, T, V, Lmax, int(blank), _p(nll, C.c_double),
                            _p(grad, C.c_double))
    return nll, grad


def pg_ctc_step(logits, targets, in_len=None, tgt_len=None, uniforms=None, seed=0, K=16, blank=0,
                reward_mode=0, baseline_mode=1, baseline_value=0.0, w_pg=1.0, w_ctc=1.0,
                want_grad=True):
    """The whole CPU path in one call (what bench.py times as the CPU arm)."""
    logits = _f32(logits)
    B, T, V = logits.shape
    targets = _i32(targets)
    Lmax = targets.shape[1]
    in_len, tgt_len, uniforms = _i32(in_len), _i32(tgt_len), _f32(uniforms)
    if uniforms is not None:
        K = uniforms.shape[1]
    rewards = np.zeros((B, K), np.float32)
    nll = np.zeros(B, np.float64)
    dlogits = np.zeros((B, T, V), np.float32) if want_grad else None
    loss = lib().orc_pg_ctc_step(_p(logits, C.c_float), _p(targets, C.c_int32),
                                 _p(in_len, C.c_int32), _p(tgt_len, C.c_int32),
                                 _p(uniforms, C.c_float), int(seed), B, T, V, K, Lmax, int(blank),
                                 int(reward_mode), int(baseline_mode), float(baseline_value),
                                 float(w_pg), float(w_ctc), _p(rewards, C.c_float),
                                 _p(nll, C.c_double), _p(dlogits, C.c_float))
    return loss, rewards, nll, dlogits
