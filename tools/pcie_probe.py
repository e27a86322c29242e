"""PCIe and host-side overhead probe for the HostPipeline e2e path (what bounds e2e once the kernel is short)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pgasr_b200  # noqa: E402
from tests.synth import make_batch  # noqa: E402

dev = torch.device("cuda:0")
B, T, V, K, L = 64, 500, 30, 16, 100
n = B * T * V
h_in = [torch.randn(n).pin_memory() for _ in range(4)]
h_out = [torch.empty(n).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, device=dev) for _ in range(4)]
d_out = [torch.randn(n, device=dev) for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
N = 400


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6


def h2d():
    with torch.cuda.stream(s1):
        for i in range(N):
            d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        for i in range(N):
            h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)


def both():
    for i in range(N):
        with torch.cuda.stream(s1):
            d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)


mb = n * 4 / 1e6
for name, fn in (("H2D only", h2d), ("D2H only", d2h), ("H2D + D2H concurrently", both)):
    fn()
    us = timed(fn)
    print(f"{name:26s} {us:7.1f} us per {mb:.2f} MB copy  -> {mb / us * 1e3:6.1f} GB/s per direction")

# pipeline: submit cost on the host, and throughput at several depths
lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=1)
pin = lambda a: torch.from_numpy(a).pin_memory()
hl, ht, hil, htl = pin(lg), pin(tg), pin(il), pin(tl)
for depth in (1, 2, 3, 4, 6):
    pipe = pgasr_b200.HostPipeline(B, T, V, K, L, depth=depth)
    outs = [pipe.output_buffers() for _ in range(depth)]
    for i in range(10):
        pipe.submit(hl, ht, hil, htl, out=outs[i % depth], seed=i)
    pipe.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sub = 0.0
    for i in range(N):
        a = time.perf_counter()
        pipe.submit(hl, ht, hil, htl, out=outs[i % depth], seed=i)
        sub += time.perf_counter() - a
    pipe.wait()
    dt = time.perf_counter() - t0
    print(f"depth {depth}: {dt / N * 1e6:7.1f} us/step  ({B * N / dt:9.0f} utt/s), host time inside submit {sub / N * 1e6:6.1f} us/step")
    pipe.close()

# can one direction go faster with the copy split over several streams (several copy engines)?
for parts in (2, 4):
    ss = [torch.cuda.Stream() for _ in range(parts)]
    chunk = n // parts

    def h2d_split():
        for i in range(N):
            for j, st in enumerate(ss):
                with torch.cuda.stream(st):
                    d_in[i % 4][j * chunk:(j + 1) * chunk].copy_(h_in[i % 4][j * chunk:(j + 1) * chunk], non_blocking=True)

    def d2h_split():
        for i in range(N):
            for j, st in enumerate(ss):
                with torch.cuda.stream(st):
                    h_out[i % 4][j * chunk:(j + 1) * chunk].copy_(d_out[i % 4][j * chunk:(j + 1) * chunk], non_blocking=True)

    def both_split():
        for i in range(N):
            for j, st in enumerate(ss):
                with torch.cuda.stream(st):
                    d_in[i % 4][j * chunk:(j + 1) * chunk].copy_(h_in[i % 4][j * chunk:(j + 1) * chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)

    for name, fn in ((f"H2D in {parts} streams", h2d_split), (f"D2H in {parts} streams", d2h_split),
                     (f"H2D in {parts} streams + D2H", both_split)):
        fn()
        us = timed(fn)
        print(f"{name:26s} {us:7.1f} us per {mb:.2f} MB  -> {mb / us * 1e3:6.1f} GB/s")
