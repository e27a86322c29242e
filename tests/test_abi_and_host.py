"""CPU: the C-ABI library loads and exports what include/pgasr.h declares; host-side logic; no compute."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgasr.h")).read()
    return sorted(set(re.findall(r"PGASR_API\s+[\w\s\*]+?\b(pgasr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import pgasr_b200
    from pgasr_b200 import _native
    names = declared_symbols()
    assert len(names) >= 16
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pgasr.h but not exported"
    assert sorted(_native.SIGNATURES) == names, "ctypes table and header disagree"
    assert _native.lib().pgasr_abi_version() == 1
    assert _native.lib().pgasr_status_string(-4) == b"workspace too small"


def test_header_argument_counts_match_ctypes_table():
    from pgasr_b200 import _native
    text = open(os.path.join(ROOT, "include", "pgasr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _native.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_no_device_is_loud_not_a_fallback():
    """Without a GPU the product path raises; it never computes on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pgasr_b200
    from pgasr_b200 import functional as F
    assert pgasr_b200._native.lib().pgasr_device_check() != 0
    with pytest.raises(TypeError):
        F.softmax_sample(torch.zeros(1, 4, 5), K=2)
    with pytest.raises(RuntimeError):
        pgasr_b200.metrics.edit_dist("kitten", "sitting")
    with pytest.raises(RuntimeError):
        pgasr_b200.CTCdecoder.collapse_fn("aab")
    with pytest.raises(RuntimeError):
        pgasr_b200.CTCdecoder.CTCDecoder(None).decode(np.full((4, 3), 1 / 3))
    with pytest.raises(TypeError):
        pgasr_b200.loss.customNLLLoss()(torch.zeros(2, 2, 3), torch.zeros(2, 2, dtype=torch.long))
    with pytest.raises(pgasr_b200._native.PgasrError) as ei:      # the host-buffer API has no CPU path either
        pgasr_b200.HostPipeline(2, 20, 30, 4, 5)
    assert ei.value.status == -3


def test_workspace_queries_are_pure():
    from pgasr_b200 import _native
    L = _native.lib()
    assert L.pgasr_ctc_workspace_bytes(64, 500, 30, 100) >= 64 * 500 * 201 * 8
    assert L.pgasr_ctc_workspace_bytes(1, 10, 5, 600) == 0          # Lmax over the documented limit
    a = L.pgasr_pg_ctc_step_workspace_bytes(64, 500, 30, 16, 100)
    assert a >= 64 * 500 * 256 * 8                                   # the fp64 half-lattice of the fused kernel
    assert L.pgasr_pg_ctc_step_workspace_bytes(0 - 1, 500, 30, 16, 100) == 0


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "policy-gradient-asr_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b", src) and "oracle/pgasr_oracle.c" not in src.replace("oracle/pgasr_oracle.c", ""):
                    if re.search(r"^\s*(from|import)\s+oracle|cport|pyref|libpgasr_oracle", src, flags=re.M):
                        bad.append(f)
    assert not bad, bad


def test_dropin_modules_have_upstream_names():
    sys.path.insert(0, os.path.join(ROOT, "policy-gradient-asr_b200", "dropin"))
    try:
        for mod, names in [("metrics", ["edit_dist", "evaluate", "save_predictions"]),
                           ("CTCdecoder", ["collapse_fn", "CTCDecoder"]),
                           ("policy_grad", ["reward"]), ("loss", ["customNLLLoss"])]:
            sys.modules.pop(mod, None)
            m = __import__(mod)
            for n in names:
                assert hasattr(m, n), (mod, n)
            sys.modules.pop(mod, None)
    finally:
        sys.path.pop(0)
    import inspect
    import pgasr_b200
    assert list(inspect.signature(pgasr_b200.metrics.edit_dist).parameters) == ["s1", "s2"]
    assert list(inspect.signature(pgasr_b200.policy_grad.reward).parameters) == \
        ["true_y", "pred_y", "t", "ind2char", "ctc_decoder"]
    assert list(inspect.signature(pgasr_b200.CTCdecoder.CTCDecoder.decode).parameters) == \
        ["self", "probs", "beam_size", "blank"]
    assert inspect.signature(pgasr_b200.CTCdecoder.CTCDecoder.decode).parameters["beam_size"].default == 100
    assert inspect.signature(pgasr_b200.loss.customNLLLoss.__init__).parameters["ignore_index"].default is None


def test_save_predictions(tmp_path):
    import pgasr_b200
    pgasr_b200.metrics.save_predictions(["ab", "c d"], ["ab", "cd"], str(tmp_path))
    assert open(tmp_path / "predicted.txt").read() == "ab|ab\nc d|cd\n"
    with pytest.raises(IndexError):
        pgasr_b200.metrics.save_predictions(["a", "b"], ["a"], str(tmp_path))


def test_sharding_covers_every_utterance_once():
    from pgasr_b200.distributed import balanced_assignment, shard_range
    for n in [0, 1, 7, 64, 257]:
        for world in [1, 2, 3, 8]:
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    lens = [500, 10, 400, 20, 300, 30, 200, 40]
    parts = balanced_assignment(lens, 2)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(lens[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 100
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from pgasr_b200.distributed import MovingBaseline, allreduce_mean, allreduce_reward_stats, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    allR = -torch.arange(24, dtype=torch.float32).reshape(6, 4)       # 6 utterances x K=4
    lo, hi = shard_range(6, rank, world)
    mean, var, n = allreduce_reward_stats(allR[lo:hi])
    loss = allreduce_mean(torch.tensor(float(rank + 1)))
    mb = MovingBaseline(0.5)
    mb.update(allR[lo:hi])
    v2 = mb.update(allR[lo:hi] * 0)
    out.put((rank, mean, var, n, float(loss), v2))
    dist.destroy_process_group()


def test_gloo_world2_reward_stats():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    allR = -np.arange(24, dtype=np.float64)
    for rank, mean, var, n, loss, v2 in res:
        assert n == 24 and abs(mean - allR.mean()) < 1e-9 and abs(var - allR.var()) < 1e-9
        assert abs(loss - 1.5) < 1e-6
        assert abs(v2 - 0.5 * allR.mean()) < 1e-9


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the oracle port on the host cores and prints one JSON line."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_algorithmic_ops_match_the_survey_figures():
    """at T=500, V=30, K=16, L=100 the path is 15 k exponentials, 201 k lattice state updates over 500
    dependent frames and K*L_h*L edit-distance cells per utterance; bench.py reports them times the measured rate."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = types.SimpleNamespace(T=500, V=30, K=16, L=100, regime="random", w_pg=1.0, w_ctc=1.0)
    ops = bench.algorithmic_ops(a, 2.0)
    per = ops["per_utterance"]
    assert per["sampler_exp"] == 15000 and per["ctc_state_updates_alpha_plus_beta"] == 2 * 500 * 201
    assert per["levenshtein_cells"] == 16 * ops["hyp_len_assumed"] * 100 and per["ctc_dependent_frames"] == 500
    assert ops["achieved"]["sampler_exp_per_s"] == 30000.0 and "ctc_dependent_frames_per_s" not in ops["achieved"]
    a.regime, a.w_pg = "peaky", 0.0
    ops = bench.algorithmic_ops(a, 1.0)
    assert "levenshtein_cells" not in ops["per_utterance"] and ops["hyp_len_assumed"] is None
