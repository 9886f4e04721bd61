#!/usr/bin/env python
"""Top stall locations (SASS lines) of the first kernel in an .ncu-rep: python scripts/ncu_hot.py file.ncu-rep [N]"""
import csv
import subprocess
import sys

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si = hdr.index("Warp Stall Sampling (All Samples)")
src = hdr.index("Source")
ex = hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > si]
tot = sum(float(r[si] or 0) for r in body) or 1.0
print("total samples", tot, " instructions", len(body))
for r in sorted(body, key=lambda r: -float(r[si] or 0))[:n]:
    print(f"{float(r[si]) / tot * 100:5.1f}%  exec={r[ex]:>9s}  {r[src].strip()[:100]}")
