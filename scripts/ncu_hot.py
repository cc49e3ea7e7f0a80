#!/usr/bin/env python3
"""Top stall-sample SASS instructions of one kernel in an .ncu-rep: ncu_hot.py rep kernel-regex [skip] [topN]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    if r and r[0] == "Address":
        break  # a second table (another view) follows
    if len(r) == len(hdr):
        data.append(r)
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("kernel:", lines[0][:120]); print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:top]
stall_cols = [h for h in hdr if h.startswith("stall_") or h.startswith("Stall")]
for i in sorted(order):
    r = data[i]
    n = int(r[idx["# Samples"]] or 0)
    reasons = sorted(((int(r[idx[c]] or 0), c) for c in stall_cols if (r[idx[c]] or "0").isdigit()), reverse=True)[:2]
    print(f"{i:5d} {100*n/max(tot,1):5.1f}%  {r[idx['Source']].strip()[:70]:70s} {' '.join(f'{c}={v}' for v,c in reasons if v)}")
