#!/usr/bin/env python
"""Top stalled SASS instructions of an `ncu --page source --csv` export, with region totals.

    ncu -i X.ncu-rep --page source --csv > src.csv
    python tools/ncu_src_top.py src.csv [top=40] [lo-hi:label ...]     (address ranges in hex, e.g. 0da0-2970:integrator)
Profiling aid only.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
regions = []
for a in sys.argv[3:]:
    rng, label = a.split(":")
    lo, hi = (int(v, 16) for v in rng.split("-"))
    regions.append((lo, hi, label))
iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][iA], 16)
tot = sum(int(r[iN]) for r in data)
print(f"total samples {tot}")
for lo, hi, label in regions:
    sel = [r for r in data if lo <= int(r[iA], 16) - base <= hi]
    s = sum(int(r[iN]) for r in sel)
    e = sum(int(r[iE]) for r in sel)
    by = {h: sum(int(r[i] or 0) for r in sel) for i, h in stall_cols}
    tops = sorted(by.items(), key=lambda kv: -kv[1])[:6]
    print(f"{label:12s} samples {s:7d} ({100 * s / tot:5.1f} %)  warp-instr {e:10d}  " + "  ".join(f"{h[6:]} {100 * v / max(s, 1):.0f}%" for h, v in tops))
print()
for r in sorted(data, key=lambda r: -int(r[iN]))[:top]:
    off = int(r[iA], 16) - base
    by = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    print(f"0x{off:04x} {int(r[iN]):6d} {100 * int(r[iN]) / tot:5.2f}%  {r[iS][:60]:60s} " + " ".join(f"{h}:{v}" for v, h in by if v))
