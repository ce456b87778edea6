#!/usr/bin/env python
"""Static opcode histogram of the loops of one kernel, from `cuobjdump -sass`.

    cuobjdump -sass -fun '<mangled name>' build/inst_cstr.o | python tools/sass_hist.py [--min 40]

A loop = [target, branch] of every backward BRA.  For each loop at least `--min` instructions long the script
prints the instruction count, the FP64-pipe share and the opcode mix, plus the issue-cycle estimate
`other + 2.19 * fp64` (profiles/README.md: the FP64 pipe accepts one warp instruction per ~2.19 cycles per
sub-partition, everything else issues at one per cycle).  The loops of the warp-specialised kernels are straight
line (no inner branches besides the mbarrier spin), so the static count IS the executed count per iteration.
Profiling aid only.
"""
import collections
import re
import sys

FP64 = re.compile(r"^(DFMA|DMUL|DADD|DSETP|DMNMX|MUFU\.(RCP64H|RSQ64H)|I2F\.F64|F2I\.F64|F2F\.F64|F2F\.F32\.F64|I2F\.F64\.\w+|F2I\.\w*F64)")


def main():
    mn = 40
    if "--min" in sys.argv:
        mn = int(sys.argv[sys.argv.index("--min") + 1])
    ins = []
    for ln in sys.stdin:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if not m:
            continue
        text = m.group(2).strip()
        text = re.sub(r"^@!?U?P\d+\s+", "", text)
        ins.append((int(m.group(1), 16), text))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    print(f"{len(ins)} instructions in the kernel")
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.match(r"BRA(\.\w+)*\s+.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(2), 16)
            if tgt <= a and tgt in addr_idx:
                loops.append((addr_idx[tgt], i))
    for lo, hi in loops:
        n = hi - lo + 1
        if n < mn:
            continue
        ops = collections.Counter()
        f64 = 0
        for _, t in ins[lo:hi + 1]:
            op = t.split()[0]
            ops[op] += 1
            if FP64.match(op):
                f64 += 1
        print(f"\nloop 0x{ins[lo][0]:04x}..0x{ins[hi][0]:04x}: {n} instructions, {f64} FP64-pipe, {n - f64} other; "
              f"issue-cycle estimate {n - f64 + 2.19 * f64:.0f}")
        grp = collections.Counter()
        for op, c in ops.items():
            grp[op.split(".")[0]] += c
        print("  " + "  ".join(f"{op} {c}" for op, c in grp.most_common()))


if __name__ == "__main__":
    main()
