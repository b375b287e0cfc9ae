"""Instruction mix / hot spots from `ncu -i X.ncu-rep --page source --csv -k regex:NAME > file.csv`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
ops, samp = collections.Counter(), collections.Counter()
tot = 0
hot = []
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iE or not r[iE].isdigit():
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    e, n = int(r[iE]), int(r[iN])
    ops[op] += e
    samp[op] += n
    tot += e
    hot.append((n, e, r[iS].strip()))
print("total warp-instructions", tot, "per step", tot / steps)
stot = sum(samp.values())
for op, c in ops.most_common(22):
    print(op.ljust(10), f"{c / steps:10.1f}/step {c / tot * 100:5.1f}%  samples {samp[op] / max(stot, 1) * 100:5.1f}%")
print("--- top sampled instructions")
for n, e, src in sorted(hot, reverse=True)[:25]:
    print(f"{n / max(stot, 1) * 100:5.1f}%  exec/step {e / steps:8.1f}  {src[:90]}")
