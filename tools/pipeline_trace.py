"""Timeline of the HostPipeline stages (timing build: python policy-gradient-asr_b200/build.py --timing;
PGASR_LIB=policy-gradient-asr_b200/lib/libpgasr_b200_timing.so python tools/pipeline_trace.py)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pgasr_b200  # noqa: E402
from tests.synth import make_batch  # noqa: E402

B, T, V, K, L = 64, 500, 30, 16, 100
depth = int(os.environ.get("DEPTH", "4"))
lg, tg, il, tl, _ = make_batch(B, T, V, K, L, seed=1)
pin = lambda a: torch.from_numpy(a).pin_memory()
hl, ht, hil, htl = pin(lg), pin(tg), pin(il), pin(tl)
pipe = pgasr_b200.HostPipeline(B, T, V, K, L, depth=depth)
outs = [pipe.output_buffers() for _ in range(depth)]
lib = pgasr_b200._native.lib()
lib.pgasr_host_debug_times.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
N = int(os.environ.get("N", "400"))
for i in range(N):
    pipe.submit(hl, ht, hil, htl, out=outs[i % depth], seed=i)
pipe.wait()
rows = []
for slot in range(depth):
    ms = (ctypes.c_float * 6)()
    assert lib.pgasr_host_debug_times(pipe._h, slot, ms) == 0
    rows.append(list(ms))
rows.sort()
t0 = rows[0][0]
print("step   H2D start   H2D end   K start   K end   D2H start   D2H end   (us, relative)")
for r in rows:
    print("      " + "  ".join(f"{(x - t0) * 1e3:9.1f}" for x in r))
print("durations (us): " + "   ".join(f"H2D {1e3*(r[1]-r[0]):5.1f} K {1e3*(r[3]-r[2]):5.1f} D2H {1e3*(r[5]-r[4]):5.1f}" for r in rows))
print(f"cadence: {1e3 * (rows[-1][0] - rows[0][0]) / (len(rows) - 1):6.1f} us/step")
