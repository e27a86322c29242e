#!/usr/bin/env python
"""SURVEY 8f.4 / BASELINE.json configs[3]: a full PG training step -- acoustic model forward/backward (stock torch
modules: the upstream Encoder's layer sizes, model.py:34-56, plus the per-frame Linear(512, V) head upstream never
had) + the fused PG+CTC loss of this repository in the `criterion(model_out, t)` slot (model.py:235-238) -- data
parallel over the GPUs of one box, one process per GPU, NCCL all-reduce of the model gradients (DistributedDataParallel).
The loss needs no collective: every rank scores its own utterances.

    python examples/acoustic_harness.py --global-batch 256 --steps 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        examples/acoustic_harness.py --global-batch 256 --steps 20

This is a harness, not a product path: the encoder is cuDNN's LSTM, features and transcripts are synthetic.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as tF  # noqa: E402

import pgasr_b200  # noqa: E402


class AcousticModel(nn.Module):
    """Instance-normalised 120-dim features -> Linear(120,512) -> 3 x BLSTM(256) -> Linear(512,V) logits [B,T,V]."""

    def __init__(self, V, feat=120):
        super().__init__()
        self.inp = nn.Linear(feat, 512)
        self.drop = nn.Dropout(0.5)
        self.blstm = nn.LSTM(512, 256, num_layers=3, dropout=0.3, bidirectional=True, batch_first=True)
        self.head = nn.Linear(512, V)

    def forward(self, x, lengths):
        # x [B, feat, T]; normalise every feature row over the utterance (upstream: InstanceNorm2d on [B,1,feat,T])
        x = (x - x.mean(-1, keepdim=True)) / (x.std(-1, keepdim=True) + 1e-5)
        h = self.drop(tF.leaky_relu(self.inp(x.transpose(1, 2))))
        packed = nn.utils.rnn.pack_padded_sequence(h, lengths.cpu(), batch_first=True, enforce_sorted=False)
        out, _ = self.blstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=x.shape[-1])
        return self.head(out)


def run(D, global_batch=256, T=500, V=30, K=16, L=100, steps=20, warmup=3):
    """One data-parallel training job on the process group of `D` (bench.Dist: world, rank, local, dev).  Returns the
    per-step time (max over ranks) and its breakdown: the loss step alone, the step without the gradient all-reduce
    (DDP no_sync) and the all-reduce of a gradient-sized flat buffer alone."""
    world, rank, dev = D.world, D.rank, D.dev
    lo, hi = pgasr_b200.distributed.shard_range(global_batch, rank, world)
    B = hi - lo
    torch.manual_seed(1234 + rank)
    model = AcousticModel(V).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[D.local])
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)                   # model.py:207
    crit = pgasr_b200.PolicyGradCTCLoss(K=K, reward="cer", baseline="mean", seed=rank)
    feats = torch.randn(B, 120, T, device=dev)
    flen = torch.full((B,), T, dtype=torch.int32, device=dev)
    trans = torch.randint(1, V, (B, L), dtype=torch.int32, device=dev)
    tlen = torch.full((B,), L, dtype=torch.int32, device=dev)

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        if world > 1 and not sync:
            with model.no_sync():
                loss = crit(model(feats, flen), trans, flen, tlen)
                loss.backward()
        else:
            loss = crit(model(feats, flen), trans, flen, tlen)
            loss.backward()
        opt.step()
        return loss

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return D.max_over_ranks(e0.elapsed_time(e1) / n), out

    for _ in range(warmup):
        step()
    ms, loss = timed(step, steps)
    ms_nosync, _ = timed(lambda: step(sync=False), max(steps // 2, 3)) if world > 1 else (ms, None)
    # the gradient all-reduce alone: one flat fp32 buffer of the model's size (what DDP's buckets add up to)
    ar_ms = 0.0
    if world > 1:
        flat = torch.zeros(n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        ar_ms, _ = timed(lambda: dist.all_reduce(flat), 20)
        loss = pgasr_b200.distributed.allreduce_mean(loss)
    # the loss step alone, on this rank's logits
    with torch.no_grad():
        logits = model(feats, flen).detach()
    for _ in range(3):
        crit(logits.requires_grad_(True), trans, flen, tlen)
    n0 = pgasr_b200._native.lib().pgasr_launch_count()
    loss_ms, _ = timed(lambda: crit(logits.requires_grad_(True), trans, flen, tlen), 20)
    per_step = (pgasr_b200._native.lib().pgasr_launch_count() - n0) / 20
    return {"what": "full PG training step: acoustic model fwd/bwd + fused PG+CTC loss + optimizer",
            "n_gpus": world, "global_batch": global_batch, "B_per_gpu": B, "T": T, "V": V, "K": K, "L": L,
            "ms_per_step": float(ms), "utt_per_s": global_batch / float(ms) * 1e3,
            "ms_per_step_without_allreduce": float(ms_nosync), "allreduce_exposed_ms": float(ms - ms_nosync),
            "allreduce_alone_ms": float(ar_ms), "allreduce_bytes": 4 * n_params,
            "loss_step_ms": float(loss_ms), "loss_launches_per_step": per_step, "loss": float(loss),
            "mean_reward": float(crit.last["rewards"].mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--T", type=int, default=500)
    ap.add_argument("--V", type=int, default=30)
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    D = bench.Dist()
    res = run(D, args.global_batch, args.T, args.V, args.K, args.L, args.steps, args.warmup)
    if D.rank == 0:
        print(json.dumps(res))
    D.close()


if __name__ == "__main__":
    main()
