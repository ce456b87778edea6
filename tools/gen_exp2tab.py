#!/usr/bin/env python
"""Writes pcgym_b200/csrc/exp2tab.inc: 2^(j/256) for j = 0..255, correctly rounded to fp64 (60-digit decimal)."""
import os
from decimal import Decimal, getcontext

getcontext().prec = 60
vals = [float(Decimal(2) ** (Decimal(j) / Decimal(256))) for j in range(256)]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pcgym_b200", "csrc", "exp2tab.inc")
with open(out, "w") as f:
    f.write("// 2^(j/256), j = 0..255, correctly rounded (tools/gen_exp2tab.py)\n")
    for i in range(0, 256, 4):
        f.write("    " + ", ".join(repr(v) for v in vals[i:i + 4]) + ",\n")
