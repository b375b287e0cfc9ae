"""Per-kernel registers / stack / static shared memory of the built library (cuobjdump -res-usage),
sorted by stack use: the spill check to read before spending GPU time.  Usage:
python tools/resource_usage.py > profiles/rNN_resource_usage.txt"""
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "keypoint_moseq_b200", "libkpms_b200.so")
txt = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.split("\n")
rows, name = [], None
for line in txt:
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        rows.append((name,) + tuple(int(x) for x in m.groups()))
        name = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -res-usage libkpms_b200.so (sm_100a), one line per kernel instantiation")
print("# STACK > 0 marks spills or local arrays; dynamic shared memory is set at launch and not listed")
print("#  REG  STACK  SHARED  LOCAL  kernel")
for (n, reg, st, sh, lo), dn in sorted(zip(rows, names), key=lambda a: (-a[0][2], -a[0][1])):
    dn = re.sub(r"^void ", "", dn)
    dn = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", dn)
    print(f"{reg:5d} {st:6d} {sh:7d} {lo:6d}  {dn[:140]}")
