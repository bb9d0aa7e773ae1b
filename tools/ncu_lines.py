"""Per-CUDA-source-line executed warp instructions + stall samples from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` (line rows carry the per-line aggregates)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
fname, hdr, out = "", None, []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    si, ei = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    try:
        out.append((int(r[ei]), int(r[si]), fname, int(r[0]), r[1].strip()))
    except ValueError:
        pass
tot, ts = sum(o[0] for o in out), max(sum(o[1] for o in out), 1)
print("total warp instructions", tot, "stall samples", ts)
for n, s, f, ln, text in sorted(out, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{f}:{ln:<5d} {100 * n / tot:5.1f}% stall {100 * s / ts:5.1f}%  {text[:100]}")
