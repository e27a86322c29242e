"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows, in address order) with
`nvdisasm -g -c` of the same cubin (SASS with //## File ... line N markers) by instruction offset.

    python tools/ncu_lines.py <report.ncu-rep> <mangled kernel name substring> [top N]

Needs the library built from the same sources as the profiled run (policy-gradient-asr_b200/lib/libpgasr_b200.so).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "policy-gradient-asr_b200", "lib", "libpgasr_b200.so")


def disasm(kernel_sub):
    """[(offset, file, line, text)] of the first kernel whose mangled name contains kernel_sub."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or "-" in f.split(".")[0]:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", f], cwd=tmp, capture_output=True, text=True).stdout
        out, cur, on, fl = [], None, False, ("?", 0)
        for line in txt.splitlines():
            m = re.match(r"\.text\.(\S+):", line)
            if m:
                if on and out:
                    return out
                on = kernel_sub in m.group(1)
                continue
            if not on:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                fl = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), fl[0], fl[1], m.group(2).strip()))
        if on and out:
            return out
    raise SystemExit("kernel not found in " + LIB)


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sass = disasm(ksub)
    by_off = {o: (f, l, t) for o, f, l, t in sass}
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[2][iA], 16)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
    tot = 0
    for r in rows[2:]:
        if len(r) <= iE:
            continue
        off = int(r[iA], 16) - base
        f, l, _ = by_off.get(off, ("?", 0, ""))
        n = int(r[iN] or 0)
        a = agg[(f, l)]
        a[0] += n
        a[1] += int(r[iE] or 0)
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[2][h[6:]] += v
        op = r[iS].strip().split()
        op = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        a[3][op] += n
        tot += n
    print(f"# {os.path.basename(rep)} / {ksub}: {tot} warp-stall samples; top {top} source lines")
    print(f"{'file:line':28s} {'samples':>8s} {'share':>6s} {'warp-inst':>10s}  top stall reasons | top opcodes by samples")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = " ".join(f"{k}={v}" for k, v in a[2].most_common(3))
        ops = " ".join(f"{k}={v}" for k, v in a[3].most_common(3))
        print(f"{f + ':' + str(l):28s} {a[0]:8d} {100 * a[0] / max(tot, 1):5.1f}% {a[1]:10d}  {st} | {ops}")


if __name__ == "__main__":
    main()
