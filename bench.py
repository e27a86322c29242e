#!/usr/bin/env python
"""bench.py -- PG-loss + CTC utterances/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    the CPU arm: the oracle port on the host cores

A step is one pass of the hot path (sample K hypotheses -> collapse -> edit distance -> reward -> baseline ->
REINFORCE gradient, plus CTC alpha-beta loss and gradient; loss scalar + dlogits out) over one batch of B
synthetic utterances per GPU at BASELINE.json configs[1]'s shape: B=64, T=500, V=30, K=16, label length 100.
Utterances are independent, so N GPUs run N batches with no data-path collective (scaling: weak).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--T", type=int, default=500)
    ap.add_argument("--V", type=int, default=30)
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--regime", default="random", choices=["random", "peaky"])
    ap.add_argument("--cpu-batch", type=int, default=0, help="utterances per CPU step (0: same as --batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def algorithmic_bytes_per_utt(T, V, L, K):
    # read logits + write dlogits + targets + lengths + nll, R[K], logp[K]
    return 8 * T * V + 4 * L + 8 * K + 12


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def traffic_from_profile(workload_key):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if d.get("workload") == workload_key:
                return d.get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(args, B):
    from oracle import cport
    from tests.synth import make_batch
    cport.set_num_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses all cores
    logits, targets, in_len, tgt_len, _ = make_batch(B, args.T, args.V, args.K, args.L, seed=1234, regime=args.regime)

    def step(i):
        return cport.pg_ctc_step(logits, targets, in_len, tgt_len, None, seed=0x5EED + i, K=args.K)
    return step, cport.max_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.cpu_batch or args.batch
    step, cores = cpu_step_fn(args, B)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = f"{args.steps} steps x {B} utterances (T={args.T},V={args.V},K={args.K},L={args.L}), C oracle port, OpenMP over utterances"
    line = {
        "impl": "reference", "metric": "pg_ctc_loss_utterances_per_sec", "value": val, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, B, note="CPU arm; upstream is pure Python (cannot travel to the GPU box), so this is the C port of it"),
        "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, B, note=None):
    cfg = {"workload": f"configs[1]: PG loss (sample/collapse/edit-distance/reward/baseline/gradient) + CTC, "
                       f"B={B}/GPU, T={args.T}, V={args.V}, K={args.K}, label_len={args.L}, {args.regime} logits",
           "B_per_gpu": B, "T": args.T, "V": args.V, "K": args.K, "L": args.L, "regime": args.regime,
           "reward": "ed", "baseline": "mean", "pg_weight": 1.0, "ctc_weight": 1.0, "rng": "philox4x32-10"}
    if note:
        cfg["note"] = note
    return cfg


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, nm in names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import pgasr_b200
    from pgasr_b200 import _native, functional as F
    from tests.synth import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(_native.LIB_PATH):             # the library normally travels with the snapshot; else build it here
        if rank == 0:
            import __graft_entry__
            __graft_entry__.build()
        if world > 1:
            dist.barrier()
    _native.lib()
    assert _native.lib().pgasr_device_check() == 0, "not an sm_100 device"

    B, T, V, K, L = args.batch, args.T, args.V, args.K, args.L
    bytes_logits = B * T * V * 4
    pool = max(2, -(-int(1.25 * L2_BYTES) // bytes_logits))        # rotating inputs: pool footprint > L2
    host, devb = [], []
    for i in range(pool):
        lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=1000 * rank + i, regime=args.regime)
        h = {"logits": torch.from_numpy(lg).pin_memory(), "targets": torch.from_numpy(tg).pin_memory(),
             "in_len": torch.from_numpy(il).pin_memory(), "tgt_len": torch.from_numpy(tl).pin_memory()}
        host.append(h)
        devb.append({k: v.to(dev) for k, v in h.items()})
    ws = F.StepWorkspace(B, T, V, K, L, dev)
    want = ("rewards", "nll")

    def step(i, batch):
        return F.pg_ctc_step(batch["logits"], batch["targets"], batch["in_len"], batch["tgt_len"], K=K,
                             seed=0x5EED + i, workspace=ws, want=want)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return x

    # ---- device-resident throughput ("value") --------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(i, devb[i % pool])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _native.lib().pgasr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i, devb[i % pool])
    e1.record()
    sampler.sample()
    torch.cuda.synchronize()
    sampler.sample()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _native.lib().pgasr_launch_count() - launches0
    barrier()
    sampler.stop_flag = True
    sampler.join(1.0)
    value = world * B * args.steps / (ms * 1e-3)

    # ---- dominant kernel (roofline): the step IS one launch of pg_ctc_fused_kernel; each launch is bracketed by its
    # own CUDA event pair on the launching stream (torch's current stream) and the durations are averaged ----------
    nrep = min(args.steps, 200)
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nrep)]
    launches1 = _native.lib().pgasr_launch_count()
    for i, (k0, k1) in enumerate(pairs):
        b = devb[i % pool]
        k0.record()
        step(i, b)
        k1.record()
    torch.cuda.synchronize()
    per_step_launches = (_native.lib().pgasr_launch_count() - launches1) / nrep
    kernel_ms = sum(k0.elapsed_time(k1) for k0, k1 in pairs) / nrep
    peak, peak_src = measured_peak()
    step_bytes = B * algorithmic_bytes_per_utt(T, V, L, K)       # 120 540 B/utt at the headline shape
    achieved = step_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "pg_ctc_fused_kernel<8,512,tile,tile> (the whole step: 1 launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peak_src == "measured" else peak_src,
                "traffic": traffic_from_profile(f"B={B},T={T},V={V},K={K},L={L}"),
                "algorithmic_bytes_per_launch": step_bytes, "kernel_ms": kernel_ms,
                "launches_per_step": per_step_launches,
                "step_achieved_GBps": step_bytes * args.steps / (ms * 1e-3) / 1e9,
                "note": "latency bound, not bandwidth bound: T dependent lattice frames per utterance and only "
                        "2B=128 CTAs of work; see DESIGN.md section 5"}

    # ---- end to end through the public host-buffer API ("e2e"): pinned HOST inputs and outputs, every step copies
    # its logits/targets/lengths H2D and its loss, rewards, nll AND the full dlogits D2H inside the timed region;
    # HostPipeline keeps `depth` steps in flight so the copies of neighbouring steps overlap the kernel ----------
    e2e = None
    if not args.no_e2e:
        depth = 4
        pipe = pgasr_b200.HostPipeline(B, T, V, K, L, depth=depth)
        outs = [pipe.output_buffers() for _ in range(depth)]

        def e2e_step(i):
            h = host[i % pool]
            return pipe.submit(h["logits"], h["targets"], h["in_len"], h["tgt_len"], out=outs[i % depth], seed=0x5EED + i)
        for i in range(max(args.warmup, 3)):
            e2e_step(i)
        pipe.wait()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        pipe.wait()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        # single-step latency (submit + wait, nothing else in flight)
        lat = []
        for i in range(20):
            t1 = time.perf_counter()
            pipe.wait(e2e_step(i))
            lat.append(time.perf_counter() - t1)
        pipe.close()
        h2d = bytes_logits + B * L * 4 + 2 * B * 4
        d2h = bytes_logits + (B * K + B + 4) * 4
        e2e = {"value": world * B * args.steps / dt, "unit": "utt/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
               "api": f"HostPipeline.submit/wait (pgasr_host_* C ABI), depth {depth}, pinned host buffers",
               "outputs_to_host": "loss, rewards[B,K], nll[B], dlogits[B,T,V]",
               "sync_step_latency_ms": 1e3 * statistics.median(lat)}

    # ---- CPU baseline on rank 0 at N=1 ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = args.cpu_batch or B
        cstep, cores = cpu_step_fn(args, Bc)
        t0 = time.perf_counter()
        cstep(0)
        first = time.perf_counter() - t0
        reps = int(min(max(round(10.0 / max(first, 1e-3)), 1), 20))
        t0 = time.perf_counter()
        for i in range(reps):
            cstep(i + 1)
        dt = time.perf_counter() - t0
        cpu = {"value": Bc * reps / dt, "unit": "utt/s", "cores": cores, "kind": "port",
               "sample": f"{reps} steps x {Bc} utterances of the same workload, C oracle port (OpenMP over utterances)"}

    if rank == 0:
        line = {
            "metric": "pg_ctc_loss_utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 I/O and sampler, f64 CTC lattice, integer collapse/edit distance", "data": "synthetic",
            "config": dict(workload_config(args, B),
                           l2=f"inputs rotate over a pool of {pool} distinct batches ({pool * bytes_logits / 2**20:.0f} MiB of logits > 126 MiB L2)"),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.result(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; see main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # stdout must carry exactly one JSON line: libraries that print there (NCCL's version banner under torchrun) are
    # sent to stderr by pointing fd 1 at fd 2 for the lifetime of the run; the JSON line is written to the saved fd
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
