"""Measured DRAM traffic of every kernel of the stress grid (BASELINE.json configs[4]) and configs[2]/[1]: runs
`tools/sweep.py --one B T V K L` under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
(one GPU) and writes one JSON line per case to profiles/r02_sweep_ncu.jsonl:
  {"B":..,"T":..,"V":..,"K":..,"L":.., "dram_bytes": {kernel-name fragment: bytes of its last launch}, "ncu_us": {...}}
bench.py --config 4 divides these bytes by its own CUDA-event times (numbers taken under ncu are never bench values).

    python tools/sweep_ncu.py [out.jsonl]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sweep_ncu.jsonl")
FRAGS = ["softmax_sample_kernel", "collapse_u8_kernel", "myers_u8_kernel", "pg_advantages_kernel", "pg_grad_kernel",
         "pg_ctc_fused_kernel", "ctc_kernel", "finalize_loss_kernel"]
CASES = [(128, 1000, 30, 16, 200), (64, 500, 30, 16, 100)] + \
        [(32, T, 30, K, L) for K in (4, 16, 64) for T, L in ((250, 50), (1000, 200), (2000, 400))]

rows = []
for B, T, V, K, L in CASES:
    log = f"/tmp/sweep_ncu_{B}_{T}_{K}.csv"
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "--one", str(B), str(T), str(V), str(K), str(L)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(log):
        print("ncu failed for", (B, T, V, K, L), r.stderr[-300:], file=sys.stderr)
        continue
    lines = [l for l in open(log) if l.startswith('"')]
    rd = list(csv.DictReader(lines))
    per = {}                                            # launch id -> {metric: value}
    for row in rd:
        key = (row["ID"], row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        name = row["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        per.setdefault(key, {})[name] = v
    dram, us = {}, {}
    for (lid, kname), m in sorted(per.items(), key=lambda kv: int(kv[0][0])):
        for f in FRAGS:
            if f in kname and not (f == "ctc_kernel" and "fused" in kname):
                # the LAST launch of a kernel wins (sweep.py --one runs a warm-up first); the fused kernel appears for the
                # CTC-only call and for the whole step: keep the larger (whole step) figure under its own key
                tot = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
                dram[f] = tot
                us[f] = m.get("gpu__time_duration.sum", 0.0)
    rows.append({"B": B, "T": T, "V": V, "K": K, "L": L, "dram_bytes": dram, "ncu_us": us})
    print(json.dumps(rows[-1]), flush=True)
with open(OUT, "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
