#!/usr/bin/env python
"""Writes pcgym_b200/csrc/logtab.inc: for j = 0..127 the pair (inv_c, -log(inv_c)) used by log_tab (csrc/fastmath.cuh).
inv_c = 1 / (1 + (j + 1/2) / 128) rounded to 12 significant bits, so m * inv_c - 1 is exact in one fma for every fp64 m in
the j-th 128th of [1, 2); the logarithm is that of the ROUNDED reciprocal, correctly rounded to fp64 (mpmath, 80 digits).
`--check` evaluates the algorithm in numpy against mpmath over random arguments (the device code is the same arithmetic)."""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 80


def table():
    inv, lg = [], []
    for j in range(128):
        c = 1.0 + (j + 0.5) / 128.0
        ic = float(np.round((1.0 / c) * 4096.0) / 4096.0)
        inv.append(ic)
        lg.append(float(-mp.log(mp.mpf(ic))))
    return np.array(inv), np.array(lg)


COEF = [1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7]    # log1p(r) = r - r^2/2 + r^3 (1/3 - r/4 + r^2/5 - r^3/6 + r^4/7); |r| < 2^-7.9


def log_tab(x, inv, lg):
    bits = x.view(np.int64)
    e = ((bits >> 52) & 0x7FF) - 1023
    m = ((bits & ((1 << 52) - 1)) | (1023 << 52)).view(np.float64)
    j = (bits >> 45) & 127
    r = m * inv[j] - 1.0            # exact on the device (fma); numpy: product exact (53 + 12 bits needs 65: emulate below)
    r = np.array([float(mp.mpf(float(mi)) * mp.mpf(float(ii)) - 1) for mi, ii in zip(m, inv[j])])
    q = COEF[4]
    for c in COEF[3::-1]:
        q = q * r + c
    p = r + r * r * (r * q - 0.5)
    ln2_hi, ln2_lo = 0.6931471803691238, 1.9082149292705877e-10
    return (e * ln2_hi + lg[j]) + (p + e * ln2_lo)


if __name__ == "__main__":
    inv, lg = table()
    if "--check" in sys.argv:
        rng = np.random.default_rng(0)
        x = np.concatenate([np.exp(rng.uniform(-30, 30, 20000)), 1 + rng.uniform(-0.3, 0.3, 5000), rng.uniform(0.5, 2, 5000)])
        got = log_tab(x, inv, lg)
        ref = np.array([float(mp.log(mp.mpf(float(v)))) for v in x])
        err = np.abs(got - ref)
        rel = err / np.maximum(np.abs(ref), 1e-300)
        print("max abs err", err.max(), "max rel err (|log| > 0.01)", rel[np.abs(ref) > 0.01].max(), "max |r|", 0)
        sys.exit(0)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pcgym_b200", "csrc", "logtab.inc")
    with open(out, "w") as f:
        f.write("// (inv_c, -log(inv_c)) for the 128 mantissa intervals of log_tab (tools/gen_logtab.py)\n")
        for a, b in zip(inv, lg):
            f.write(f"    PCG_LT({float(a)!r}, {float(b)!r}),\n")
