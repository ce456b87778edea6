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


@pytest.mark.parametrize("name", ["dopri5", "tsit5"])
def test_oracle_error_weights_match_the_kernel_tableaus(name):
    from oracle.integrate import TABLEAUS
    _, E = parse({"dopri5": "Dopri5", "tsit5": "Tsit5"}[name])
    np.testing.assert_array_equal(np.array(TABLEAUS[name]["E"]), E)


def parse_rodas4():
    src = open(os.path.join(ROOT, "pcgym_b200", "csrc", "tableaus.cuh")).read()
    body = src[src.index("struct Rodas4"):]
    num = lambda s: float(eval(s.strip(), {"__builtins__": {}}))
    out = {}
    for key, pat in (("A", r"A\[\d\]\[\d\] = \{(.*?)\};"), ("C", r"C\[\d\]\[\d\] = \{(.*?)\};")):
        rows = re.findall(r"\{([^{}]*)\}", re.search(pat, body, re.S).group(1))
        out[key] = np.array([[num(v) for v in r.split(",") if v.strip()] for r in rows])
    out["gamma"] = float(re.search(r"gamma = ([0-9.eE+-]+);", body).group(1))
    return out


def test_rodas4_coefficients_order_and_l_stability():
    """The Rosenbrock tableau compiled into the kernels (scheme "rodas4") equals the oracle's restatement; with an exact
    Jacobian it converges with order 4 on a stiff nonlinear problem (embedded solution: order 3), it is stiffly accurate
    (R(inf) = 0) and exact for y' = const."""
    from oracle.integrate import RODAS4, rodas4_step
    from scipy.integrate import solve_ivp
    hdr = parse_rodas4()
    assert hdr["gamma"] == RODAS4["gamma"]
    for i in range(6):
        np.testing.assert_array_equal(hdr["A"][i][:i], RODAS4["A"][i])
        np.testing.assert_array_equal(hdr["C"][i][:i], RODAS4["C"][i])
        assert not hdr["A"][i][i:].any() and not hdr["C"][i][i:].any()

    def f(y, u):
        return np.array([-2 * y[0] + y[1] ** 2, -50 * (y[1] - np.cos(y[0]))])

    def J(y):
        return np.array([[-2, 2 * y[1]], [-50 * np.sin(y[0]), -50.0]])
    y0, T = np.array([1.0, 0.5]), 1.0
    ref = solve_ivp(lambda t, y: f(y, None), [0, T], y0, method="Radau", rtol=1e-13, atol=1e-14).y[:, -1]
    errs, embs = [], []
    for n in (20, 40, 80, 160):
        y = y0.copy()
        for _ in range(n):
            y_new, est = rodas4_step(f, J(y), y, None, T / n)
            y = y_new
        errs.append(np.abs(y - ref).max())
        embs.append(np.abs(est).max())                     # |y1 - y1_embedded| of the last step ~ h^4 (local, order 3 + 1)
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert (orders > 3.9).all() and (orders < 5.0).all(), orders
    emb_orders = np.log2(np.array(embs[:-1]) / np.array(embs[1:]))
    assert (emb_orders > 3.0).all() and emb_orders[-1] > 3.6 and (emb_orders < 4.6).all(), emb_orders   # -> 4 from below
    # y' = lambda y with lambda h -> -inf: stiffly accurate, the numerical solution goes to zero
    lam = -1e9
    y1, _ = rodas4_step(lambda y, u: lam * y, np.array([[lam]]), np.array([1.0]), None, 1.0)
    assert abs(y1[0]) < 1e-6
    y1, _ = rodas4_step(lambda y, u: np.array([3.0]), np.zeros((1, 1)), np.array([1.0]), None, 0.5)
    assert abs(y1[0] - 2.5) < 1e-14
