#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all build/inst_cstr.o && nvdisasm -gi inst_cstr.sm_100a.cubin > lines.txt
    python tools/ncu_by_line.py sass.csv lines.txt '<mangled kernel name>' [warps] [steps]

Prints, per source line (outermost non-inlined frame in rollout_kernel.cuh when available): executed
instructions per thread per step and the share of stall samples.  Profiling aid only.
"""
import collections
import csv
import re
import sys

sass_csv, lines_txt, kernel = sys.argv[1:4]
warps = float(sys.argv[4]) if len(sys.argv) > 4 else 1024
steps = float(sys.argv[5]) if len(sys.argv) > 5 else 99

# offset -> (file, line) from nvdisasm -g, restricted to the kernel's .text section
off2line, chain, fresh, inside = {}, [], True, False
for ln in open(lines_txt):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kernel
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:                      # nvdisasm -gi: innermost frame first, one line per inlining level
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1), int(m.group(2))))
        chain += [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        fresh = True
        if chain:
            off2line[int(m.group(1), 16)] = list(chain)

rows = list(csv.reader(open(sass_csv)))
hdr, data = rows[1], rows[2:]
iA, iN, iE = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][iA], 16)
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0
for r in data:
    off = int(r[iA], 16) - base
    frames = off2line.get(off, [("?", 0)])
    # attribute to the outermost frame inside the kernel file
    key = next(((f.split("/")[-1], l) for f, l in reversed(frames) if f.endswith("rollout_kernel.cuh")), (frames[-1][0].split("/")[-1], frames[-1][1]))
    k = key + ("",)
    agg[k][0] += int(r[iE]) / warps / steps
    agg[k][1] += int(r[iN])
    tot += int(r[iN])
print(f"{'file:line':34s} {'instr/thread-step':>18s} {'samples %':>10s}")
for k, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{k[0]}:{k[1]:<5d} {k[2]:14s} {e:14.1f} {100 * s / tot:10.1f}")
