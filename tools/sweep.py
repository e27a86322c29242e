"""BASELINE.json configs[2] and configs[4] on one GPU: the CTC kernel alone at B=128, T=1000, L=200, and the stress
sweep K in {4,16,64} x T in {250,1000,2000} (label length up to 400) with the achieved algorithmic GB/s of every
stand-alone kernel and of the whole step.  One JSON line per case on stdout.

    python tools/sweep.py [--quick]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pgasr_b200 import functional as F  # noqa: E402
from tests.synth import make_batch  # noqa: E402

PEAK = 6557.1
if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def case(B, T, V, K, L, regime="random", dev=None, seed=None, reps=20, quiet=False):
    dev = dev or torch.device("cuda:0")
    lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=B + T + K + L if seed is None else seed, regime=regime)
    t = lambda a: torch.from_numpy(a).to(dev)
    lg, tg, il, tl = t(lg), t(tg), t(il), t(tl)
    out = {"B": B, "T": T, "V": V, "K": K, "L": L, "regime": regime, "kernels": {}}
    tm = lambda fn: timed(fn, reps=reps)

    def rec(name, ms, nbytes):
        out["kernels"][name] = {"ms": round(ms, 4), "bytes": nbytes, "algorithmic_GBps": round(nbytes / ms / 1e6, 2),
                                "frac_of_measured_hbm": round(nbytes / ms / 1e6 / PEAK, 5)}

    # SURVEY 8(d) per-kernel algorithmic bytes
    smp, logp, probs = F.softmax_sample(lg, il, K=K, seed=1, return_probs=True)
    rec("K1 softmax_sample", tm(lambda: F.softmax_sample(lg, il, K=K, seed=1)), B * (4 * T * V + K * T + 4 * K))
    hyp, hl = F.collapse(smp, il, blank=0)
    rec("K2 collapse", tm(lambda: F.collapse(smp, il, blank=0)), B * (2 * K * T + 4 * K))
    dist = F.edit_distance(hyp, hl.reshape(-1), tg, tl, rows_per_ref=K, vocab=V)
    rec("K3 edit_distance", tm(lambda: F.edit_distance(hyp, hl.reshape(-1), tg, tl, rows_per_ref=K, vocab=V)),
        B * (K * T + 4 * L + 4 * K))
    rew, adv, terms = F.pg_advantages(dist, tl, logp, Lmax=L)
    rec("K4 pg_grad", tm(lambda: F.pg_grad(smp, adv, il, V=V, scale=1.0 / (B * K))), B * (K * T + 4 * K + 4 * T * V))
    rec("K5 ctc_loss_grad", tm(lambda: F.ctc_loss_grad(lg, tg, il, tl)), B * (8 * T * V + 4 * L + 4))
    ws = F.StepWorkspace(B, T, V, K, L, dev)
    step_bytes = B * (8 * T * V + 4 * L + 8 * K + 12)
    outs = ws.outputs(())
    ms = tm(lambda: F.pg_ctc_step(lg, tg, il, tl, K=K, seed=3, workspace=ws, out=outs))
    rec("whole step (pgasr_pg_ctc_step)", ms, step_bytes)
    out["step_utt_per_s"] = round(B / ms * 1e3, 1)
    out["ctc_utt_per_s"] = round(B / out["kernels"]["K5 ctc_loss_grad"]["ms"] * 1e3, 1)
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


# kernel-name fragments of the ncu launch list -> the row of a case they belong to
NCU_ROWS = {"softmax_sample_kernel": "K1 softmax_sample", "collapse_u8_kernel": "K2 collapse", "myers_u8_kernel": "K3 edit_distance",
            "pg_grad_kernel": "K4 pg_grad", "pg_ctc_fused_kernel": "whole step (pgasr_pg_ctc_step)"}


def merge_measured_traffic(cases, path):
    """profiles/r02_sweep_ncu.jsonl (tools/sweep_ncu.py: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch,
    one GPU) -> measured DRAM bytes and measured GB/s (those bytes / this run's CUDA-event time) per kernel."""
    if not os.path.exists(path):
        return
    rows = [json.loads(l) for l in open(path) if l.strip()]
    for c in cases:
        for r in rows:
            if all(r.get(k) == c[k] for k in ("B", "T", "V", "K", "L")):
                for frag, dram in r.get("dram_bytes", {}).items():
                    name = NCU_ROWS.get(frag)
                    if name in c["kernels"] and c["kernels"][name]["ms"] > 0:
                        k = c["kernels"][name]
                        k["dram_bytes_ncu"] = dram
                        k["measured_GBps"] = round(dram / k["ms"] / 1e6, 2)


if __name__ == "__main__":
    if "--one" in sys.argv:                           # one case, few repetitions: the driver of tools/sweep_ncu.py
        i = sys.argv.index("--one")
        B, T, V, K, L = (int(x) for x in sys.argv[i + 1:i + 6])
        case(B, T, V, K, L, reps=1, quiet=True)
        sys.exit(0)
    quick = "--quick" in sys.argv
    case(128, 1000, 30, 16, 200)                      # configs[2] shape (CTC alone: the K5 line)
    case(64, 500, 30, 16, 100)                        # configs[1] shape, kernel by kernel
    for K in ((16,) if quick else (4, 16, 64)):
        for T, L in ((250, 50), (1000, 200), (2000, 400)):
            case(32, T, 30, K, L)
    case(64, 500, 30, 16, 100, regime="peaky")
