"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with --seq, the
per-launch sequence of one evaluation (evaluations are delimited by the mcpm:axpby marker of tools/one_eval.py)."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= mv or not r[mv]:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
        out.append((r[kn], v * scale))
    return out


def short(name):
    m = re.search(r"mcpm::(\w+)", name)
    if m and "k_launch_1d" not in name:
        t = re.search(r"<([^(]*)>\(", name)
        return "mcpm:" + m.group(1) + (("<" + t.group(1) + ">") if t else "")
    m = re.search(r"mcpm::(\w+)\(.*\)::\{lambda", name)
    if m:
        return "mcpm:" + m.group(1)
    m = re.search(r"k_launch_1d<(?:.*?)mcpm::(\w+)", name)
    if m:
        return "mcpm:" + m.group(1)
    return name[:58]


def main():
    seq = load(sys.argv[1])
    names = [short(n) for n, _ in seq]
    # last evaluation only: from the last marker on
    # the last evaluation: tools/one_eval.py repeats identical evaluations, so the launch list is periodic
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 2
    first = next(i for i, n in enumerate(names) if n.startswith("mcpm:"))
    per = (len(names) - first) // reps
    start = len(names) - per
    ev = list(zip(names[start:], [t for _, t in seq[start:]]))
    tot = sum(t for _, t in ev)
    agg = OrderedDict()
    for n, t in ev:
        c, s = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, s + t)
    print(f"{len(ev)} kernels, {tot / 1e3:.2f} ms summed kernel time in the last evaluation")
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {n:58s} x{c:4d} {s / 1e3:8.3f} ms {100 * s / tot:5.1f}%  ({s / c:7.1f} us each)")
    if "--seq" in sys.argv:
        pat = sys.argv[sys.argv.index("--seq") + 1]
        print("sequence of", pat)
        print("  " + " ".join(f"{t:.0f}" for n, t in ev if pat in n))


main()
