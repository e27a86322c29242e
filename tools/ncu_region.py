"""Stall samples of one kernel by SASS region: python tools/ncu_region.py <src.csv from `ncu --page source --csv`> [lo hi]
Without lo/hi: prints sync landmarks (BAR/SYNCS/CS2R/EXIT) with running index so that regions can be picked.
With lo hi (instruction indices): per-opcode and per-stall-reason totals plus the top instructions of that range."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
iW, iWI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows[2:] if len(r) > iE]
if len(sys.argv) < 4:
    for k, r in enumerate(data):
        s = r[iS]
        if any(x in s for x in ("BAR.", "SYNCS", "EXIT", "MEMBAR", "NANOSLEEP")):
            print(k, s.strip()[:70], "samples", r[iN], "exec", r[iE])
    sys.exit(0)
lo, hi = int(sys.argv[2]), int(sys.argv[3])
tot = collections.Counter()
ops = collections.Counter()
opn = collections.Counter()
ns = ne = wf = wfi = 0
top = []
for k in range(lo, hi):
    r = data[k]
    n = int(r[iN] or 0)
    ns += n
    ne += int(r[iE] or 0)
    wf += int(r[iW] or 0)
    wfi += int(r[iWI] or 0)
    for i, h in stall:
        v = int(r[i] or 0)
        if v:
            tot[h] += v
    op = r[iS].strip().split()
    op = (op[1] if op[0].startswith("@") else op[0]).rstrip(";")
    ops[op.split(".")[0]] += n
    opn[op.split(".")[0]] += int(r[iE] or 0)
    top.append((n, k, r[iS].strip()[:60], {h: int(r[i] or 0) for i, h in stall if int(r[i] or 0)}))
print(f"range [{lo},{hi}): {ns} samples, {ne} warp-instructions, smem wavefronts {wf} (ideal {wfi})")
print(" stalls:", " ".join(f"{k}={v}" for k, v in tot.most_common()))
print(" samples by opcode:", " ".join(f"{k}={v}" for k, v in ops.most_common(14)))
print(" executed by opcode:", " ".join(f"{k}={v}" for k, v in opn.most_common(14)))
for n, k, s, d in sorted(top, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print(f"  {k:6d} {n:6d}  {s:60s} {d}")
