"""Top stall lines of one kernel from an ncu report with source info (run where the report is):
usage: python scripts/ncu_source_top.py rep.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
out = []
for r in rows:
    if hdr is None:
        if "Source" in r or "# Samples" in " ".join(r) or any("Sampling" in c for c in r):
            hdr = r
        continue
    out.append(r)
if hdr is None:
    print(txt[:3000])
    sys.exit(0)
print(hdr)
def col(name):
    for i, h in enumerate(hdr):
        if name in h:
            return i
    return None
ci = col("Warp Stall Sampling (All") or col("Samples")
si = col("Source")
print("columns:", ci, si)
def val(r):
    try:
        return float(r[ci].replace(",", ""))
    except Exception:
        return 0.0
tot = sum(val(r) for r in out)
for k, r in sorted(enumerate(out), key=lambda kr: -val(kr[1]))[:n]:
    print(f"{k:5d} {val(r):9.0f} {100 * val(r) / max(tot, 1):5.1f}%  {r[si][:150]}")
print("total samples", tot, "lines", len(out))
