#!/usr/bin/env python
"""Stall reasons of the hottest kernel in an .ncu-rep, whole kernel and for an address range:
    python tools/ncu_stalls.py rep [lo_hex hi_hex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 8 and r[0].startswith("0x")]
first = int(data[0][0], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 40
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: 0 for i in cols}
seen = set()
for r in data:
    a = int(r[0], 16) - first
    if a in seen: break
    seen.add(a)
    if lo <= a <= hi:
        for i in cols:
            tot[hdr[i]] += int(r[i] or 0)
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"{k:28s} {v:8d} {100*v/s:5.1f}%")
