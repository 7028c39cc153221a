"""Aggregate an `ncu --page source --csv` SASS listing: instruction totals by opcode and the hottest instructions."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = rows[1]
iS, iN, iI = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
ops, samp = Counter(), Counter()
tot_i = tot_s = 0
body = rows[2:]
for r in body:
    if len(r) <= iI:
        continue
    toks = r[iS].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("ATOM", "RED", "LDG", "STG", "LDS", "STS")) and "." in op else "")
    n, s = int(r[iI] or 0), int(r[iN] or 0)
    ops[op] += n
    samp[op] += s
    tot_i += n
    tot_s += s
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:12d} {100 * n / tot_i:5.1f}% inst   {100 * samp[op] / max(tot_s, 1):5.1f}% samples")
if "--top" in sys.argv:
    k = int(sys.argv[sys.argv.index("--top") + 1])
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][iN] or 0))[:k]
    print("hottest instructions (line#, samples, executed, sass)")
    for i in sorted(idx):
        print(f"  {i:5d} {int(body[i][iN]):6d} {int(body[i][iI]):9d}  {body[i][iS].strip()[:90]}")
