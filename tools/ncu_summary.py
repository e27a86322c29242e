"""Summarise ncu outputs:  launches CSV -> per-kernel mean/share;  .ncu-rep source page -> opcode histogram."""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    unit = ""
    for row in csv.DictReader(lines):
        try:
            agg[row["Kernel Name"][:70]].append(float(row["Metric Value"].replace(",", "")))
            unit = row["Metric Unit"]
        except Exception:
            pass
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: gpu__time_duration.sum per launch ({unit})")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} n={len(v):3d} mean={sum(v)/len(v):11.1f} share={sum(v)/tot:6.3f}")


def source(rep, frames):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0])
    tot = tots = 0
    for r in rows[2:]:
        if len(r) <= iE:
            continue
        parts = r[iS].strip().split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        base = op.rstrip(";").split(".")[0]
        e, s = int(r[iE] or 0), int(r[iN] or 0)
        agg[base][0] += e
        agg[base][1] += s
        tot += e
        tots += s
    print(f"# {rep}: {tot} warp instructions, {tots} samples, {tot/frames:.1f} per warp-frame")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
        print(f"{k:10s} exec={v[0]:9d} ({v[0]/frames:6.1f}/frame) samples={v[1]:6d} ({100*v[1]/max(tots,1):5.1f}%)")


def raw(rep, keys):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if any(k in h for k in keys):
            print(f"{h:75s} {u:12s} {v}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "source":
        source(sys.argv[2], float(sys.argv[3]))
    else:
        raw(sys.argv[2], sys.argv[3:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                          "launch__registers_per_thread", "sm__warps_active.avg.pct",
                                          "smsp__average_warp_latency_per_inst_issued", "smsp__inst_executed.sum",
                                          "sm__inst_executed_pipe_fp64.avg.pct", "launch__grid_size", "launch__block_size",
                                          "issue_stalled_short_scoreboard_per", "issue_stalled_wait_per",
                                          "issue_stalled_long_scoreboard_per", "issue_stalled_branch_resolving_per",
                                          "issue_stalled_barrier_per", "lts__t_bytes.sum "])
