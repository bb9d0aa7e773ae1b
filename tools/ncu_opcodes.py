"""Opcode histogram (executed warp instructions + stall samples) from `ncu -i X.ncu-rep --page source --csv`."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, ei, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
c, st = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= ei:
        continue
    toks = r[ci].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op] += int(r[ei]); st[op] += int(r[si])
tot, ts = sum(c.values()), max(sum(st.values()), 1)
print("total warp instructions", tot, "stall samples", ts)
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:24s} {n:11d} {100 * n / tot:5.1f}%   stall {100 * st[op] / ts:5.1f}%")
