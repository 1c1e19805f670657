#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libb200he.so (cuobjdump -sass): bulk-copy TMA, mbarrier, FP64 and wide integer
multiply-adds, shuffles, and local-memory spill traffic (LDL/STL), so spill regressions are visible round over round.
    python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "reference-seal-backend_b200", "lib", "libb200he.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "IMAD.WIDE", "IMAD.HI", "IMAD", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                cur[c] += 1
                break
print(f"SASS instruction counts per kernel of {os.path.basename(lib)} (sm_100a; cuobjdump -sass)\n")
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for name, c in sorted(counts.items()):
    print(f"| `{name}` | {c['total']} | " + " | ".join(str(c[k]) for k in cols) + " |")
