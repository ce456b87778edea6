"""Order conditions of the Butcher tableaus compiled into the kernels (parsed from csrc/tableaus.cuh)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(name):
    src = open(os.path.join(ROOT, "pcgym_b200", "csrc", "tableaus.cuh")).read()
    body = src[src.index(f"struct {name}"):]
    body = body[: body.index("};\n\n") + 2] if "};\n\n" in body else body
    num = lambda s: float(eval(s.strip(), {"__builtins__": {}}))
    A_txt = re.search(r"A\[\d\]\[\d\] = \{(.*?)\};", body, re.S).group(1)
    rows = re.findall(r"\{([^{}]*)\}", A_txt)
    A = np.array([[num(v) for v in r.split(",") if v.strip()] for r in rows])
    E = None
    m = re.search(r"E\[\d\] = \{(.*?)\};", body, re.S)
    if m:
        E = np.array([num(v) for v in m.group(1).split(",") if v.strip()])
    return A, E


def conditions(M, w, order):
    c = M.sum(1)
    out = [w.sum() - 1]
    if order >= 2:
        out += [w @ c - 1 / 2]
    if order >= 3:
        out += [w @ c ** 2 - 1 / 3, w @ (M @ c) - 1 / 6]
    if order >= 4:
        out += [w @ c ** 3 - 1 / 4, (w * c) @ (M @ c) - 1 / 8, w @ (M @ c ** 2) - 1 / 12, w @ (M @ (M @ c)) - 1 / 24]
    if order >= 5:
        out += [w @ c ** 4 - 1 / 5, (w * c ** 2) @ (M @ c) - 1 / 10, (w * c) @ (M @ c ** 2) - 1 / 15,
                (w * c) @ (M @ (M @ c)) - 1 / 30, w @ ((M @ c) ** 2) - 1 / 20, w @ (M @ c ** 3) - 1 / 20,
                w @ (M @ (c * (M @ c))) - 1 / 40, w @ (M @ (M @ c ** 2)) - 1 / 60, w @ (M @ (M @ (M @ c))) - 1 / 120]
    return np.abs(out).max()


def test_rk4():
    A, _ = parse("Rk4")
    M = np.zeros((4, 4)); M[:, :] = A[:4]
    assert conditions(M, A[4], 4) < 1e-15


@pytest.mark.parametrize("name", ["Dopri5", "Tsit5"])
def test_embedded_pairs(name):
    A, E = parse(name)
    M = np.zeros((7, 7)); M[:, :6] = A            # 7 stages incl. the FSAL row (= b)
    b = np.concatenate([A[6], [0.0]])
    assert conditions(M, b, 5) < 5e-14            # 5th-order solution
    assert abs(E.sum()) < 1e-15
    emb = b - E
    assert conditions(M, emb, 4) < 5e-14          # embedded 4th-order solution
    assert conditions(M, emb, 5) > 1e-5           # ... and genuinely not 5th


@pytest.mark.parametrize("name,order", [("rk4", 4), ("dopri5", 5), ("tsit5", 5)])
def test_oracle_tableaus_match_the_kernel_tableaus(name, order):
    """oracle/integrate.py restates the coefficients from the literature; they satisfy the order conditions on
    their own and equal what the kernels compile in."""
    from oracle.integrate import TABLEAUS
    tb = TABLEAUS[name]
    S = len(tb["B"])
    M = np.zeros((S, S))
    for i, row in enumerate(tb["A"]):
        M[i, : len(row)] = row
    b = np.array(tb["B"])
    assert conditions(M, b, order) < 5e-14
    A, _ = parse({"rk4": "Rk4", "dopri5": "Dopri5", "tsit5": "Tsit5"}[name])
    np.testing.assert_array_equal(A[:S, :S][np.tril_indices(S, -1)], M[np.tril_indices(S, -1)])
    np.testing.assert_array_equal(A[S][:S], b)
