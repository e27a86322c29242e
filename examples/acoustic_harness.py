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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = pgasr_b200.distributed.shard_range(args.global_batch, rank, world)
    B = hi - lo
    torch.manual_seed(1234 + rank)
    model = AcousticModel(args.V).to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)                   # model.py:207
    crit = pgasr_b200.PolicyGradCTCLoss(K=args.K, reward="cer", baseline="mean", seed=rank)
    feats = torch.randn(B, 120, args.T, device=dev)
    flen = torch.full((B,), args.T, dtype=torch.int32, device=dev)
    trans = torch.randint(1, args.V, (B, args.L), dtype=torch.int32, device=dev)
    tlen = torch.full((B,), args.L, dtype=torch.int32, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = model(feats, flen)
        loss = crit(logits, trans, flen, tlen)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        loss = pgasr_b200.distributed.allreduce_mean(loss)
    # the loss step alone, on this rank's logits
    with torch.no_grad():
        logits = model(feats, flen).detach()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        crit(logits.requires_grad_(True), trans, flen, tlen)
    k0.record()
    for _ in range(20):
        crit(logits.requires_grad_(True), trans, flen, tlen)
    k1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"what": "full PG training step: acoustic model fwd/bwd + fused PG+CTC loss + optimizer",
                          "n_gpus": world, "global_batch": args.global_batch, "B_per_gpu": B, "T": args.T, "V": args.V,
                          "K": args.K, "L": args.L, "ms_per_step": float(ms), "utt_per_s": args.global_batch / float(ms) * 1e3,
                          "loss_step_ms": k0.elapsed_time(k1) / 20, "loss": float(loss),
                          "mean_reward": float(crit.last["rewards"].mean())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
