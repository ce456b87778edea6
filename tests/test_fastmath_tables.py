"""The tables behind the branch-free device math (csrc/fastmath.cuh) are generated files: check that the committed
copies are what their generators produce and that the table-driven logarithm, restated in numpy with the exact-fma
reduction emulated in mpmath, is accurate to ~1 ulp (the device code is the same arithmetic)."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
CSRC = os.path.join(ROOT, "pcgym_b200", "csrc")


def test_log_table_matches_its_generator_and_the_algorithm_is_accurate():
    import mpmath as mp
    import gen_logtab as G
    inv, lg = G.table()
    txt = open(os.path.join(CSRC, "logtab.inc")).read()
    rows = re.findall(r"PCG_LT\(([^,]+),\s*([^)]+)\)", txt)
    assert len(rows) == 128
    got = np.array([[float(a), float(b)] for a, b in rows])
    assert np.array_equal(got[:, 0], inv) and np.array_equal(got[:, 1], lg)
    # every reciprocal has at most 12 significant bits (so m * inv_c - 1 is exact in one fma) and |r| < 2^-7.9
    assert np.all(inv * 4096 == np.round(inv * 4096))
    edges = 1.0 + np.arange(129) / 128.0
    r = np.maximum(np.abs(edges[:-1] * inv - 1), np.abs(edges[1:] * inv - 1))
    assert r.max() < 2.0 ** -7.9
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-40, 40, 1500)), 1 + rng.uniform(-0.4, 0.4, 500)])
    val = G.log_tab(x, inv, lg)
    ref = np.array([float(mp.log(mp.mpf(float(v)))) for v in x])
    assert np.max(np.abs(val - ref) / np.maximum(np.abs(ref), 1e-2)) < 4e-16 * 8      # a few ulp, absolute near x = 1


def test_exp2_table_matches_its_generator():
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    want = [float(Decimal(2) ** (Decimal(j) / Decimal(256))) for j in range(256)]
    txt = open(os.path.join(CSRC, "exp2tab.inc")).read()
    got = [float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", txt.split("\n", 1)[1])]
    assert got == want
