"""Do PCIe copies slow down while the fused kernel runs (no dependencies between them)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pgasr_b200 import functional as F
from tests.synth import make_batch
dev = torch.device("cuda:0")
B, T, V, K, L = 64, 500, 30, 16, 100
n = B * T * V
h_in = [torch.randn(n).pin_memory() for _ in range(4)]
h_out = [torch.empty(n).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, device=dev) for _ in range(4)]
d_out = [torch.randn(n, device=dev) for _ in range(4)]
lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=1)
t = lambda a: torch.from_numpy(a).to(dev)
lg, tg, il, tl = t(lg), t(tg), t(il), t(tl)
ws = F.StepWorkspace(B, T, V, K, L, dev)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
N = 600

def run(h2d, d2h, kern, wp=1.0, wc=1.0):
    torch.cuda.synchronize()
    evs = {}
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        evs[name][0].record(st)
    for i in range(N):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)
        if kern:
            with torch.cuda.stream(s3):
                F.pg_ctc_step(lg, tg, il, tl, K=K, seed=i, workspace=ws, want=(), pg_weight=wp, ctc_weight=wc)
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name][1].record(st)
    torch.cuda.synchronize()
    return {k: v[0].elapsed_time(v[1]) / N * 1e3 for k, v in evs.items()}

run(True, True, True)
for label, args in (("kernel alone", (False, False, True)), ("H2D alone", (True, False, False)), ("D2H alone", (False, True, False)),
                    ("H2D + D2H", (True, True, False)), ("H2D + kernel", (True, False, True)), ("D2H + kernel", (False, True, True)),
                    ("H2D + D2H + kernel", (True, True, True))):
    r = run(*args)
    print(f"{label:22s} us per item: " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items() if v > 1.0))
r = run(True, True, True, wp=1.0, wc=0.0); print("H2D + D2H + PG-only kernel  " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()))
r = run(True, True, True, wp=0.0, wc=1.0); print("H2D + D2H + CTC-only kernel " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()))

# same, but the kernel consumes the H2D targets and the D2H drains the kernel's own output (no event dependencies:
# racy on purpose, timing only)
def run2():
    torch.cuda.synchronize()
    evs = {}
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        evs[name][0].record(st)
    outs = [None] * 4
    for i in range(N):
        with torch.cuda.stream(s1):
            d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
        with torch.cuda.stream(s3):
            o = F.pg_ctc_step(d_in[(i + 2) % 4].view(B, T, V), tg, il, tl, K=K, seed=i, workspace=ws, want=())
            outs[i % 4] = o["dlogits"]
        if outs[(i + 2) % 4] is not None:
            with torch.cuda.stream(s2):
                h_out[i % 4].copy_(outs[(i + 2) % 4].view(-1), non_blocking=True)
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name][1].record(st)
    torch.cuda.synchronize()
    return {k: v[0].elapsed_time(v[1]) / N * 1e3 for k, v in evs.items()}
r = run2(); r = run2()
print("kernel on DMA'd buffers, D2H of kernel output: " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()))

# add the real dependencies with events, one variant at a time
def run3(dep_in, dep_out, small=False):
    torch.cuda.synchronize()
    evs = {}
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        evs[name][0].record(st)
    outs = [None] * 4
    small_h = torch.zeros(6528).pin_memory()
    small_d = torch.zeros(6528, device=dev)
    for i in range(N):
        e_in, e_k = torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(s1):
            if small:
                small_d.copy_(small_h, non_blocking=True)
            d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
            e_in.record(s1)
        with torch.cuda.stream(s3):
            if dep_in:
                s3.wait_event(e_in)
            o = F.pg_ctc_step(d_in[i % 4].view(B, T, V), tg, il, tl, K=K, seed=i, workspace=ws, want=())
            outs[i % 4] = o["dlogits"]
            e_k.record(s3)
        with torch.cuda.stream(s2):
            if dep_out:
                s2.wait_event(e_k)
            if small:
                small_h.copy_(small_d, non_blocking=True)
            h_out[i % 4].copy_(outs[i % 4].view(-1), non_blocking=True)
    for name, st in (("h2d", s1), ("d2h", s2), ("k", s3)):
        evs[name][1].record(st)
    torch.cuda.synchronize()
    return {k: v[0].elapsed_time(v[1]) / N * 1e3 for k, v in evs.items()}
for label, a in (("no deps", (False, False)), ("K waits H2D", (True, False)), ("D2H waits K", (False, True)), ("both deps", (True, True)),
                 ("both deps + small copies", (True, True, True))):
    run3(*a); r = run3(*a)
    print(f"{label:26s} " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()))
