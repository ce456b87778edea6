"""Oracle right-hand sides: numpy restatement of the six north-star models.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes
`x[nx, ...]`, `u[nu, ...]` (trailing axes = env batch, broadcast allowed) and a
parameter dict `p`, and returns `dx[nx, ...]` in float64.  `u` is the reference's
`uk` vector: manipulated inputs first, then the model's disturbance channels
when the env was configured with disturbances.

Each function cites the reference lines whose arithmetic it restates
(`/root/reference/src/pcgym/model_classes.py`).  They are pinned against the
reference executed verbatim by tests/test_oracle_rhs.py.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# metadata: names and default parameters (model_classes.py info() methods)
# --------------------------------------------------------------------------
MODELS = {
    # model_classes.py:23-43
    "cstr": dict(
        states=["Ca", "T"], inputs=["Tc"], disturbances=["Ti", "Caf"],
        params=dict(q=100.0, V=100.0, rho=1000.0, C=0.239, deltaHr=-5e4, EA_over_R=8750.0,
                    k0=7.2e10, UA=5e4, Ti=350.0, Caf=1.0),
    ),
    # model_classes.py:296-343
    "first_order_system": dict(
        states=["x"], inputs=["u"], disturbances=["None"],
        params=dict(K=1.0, tau=0.5),
    ),
    # model_classes.py:346-430
    "multistage_extraction": dict(
        states=["X1", "Y1", "X2", "Y2", "X3", "Y3", "X4", "Y4", "X5", "Y5"],
        inputs=["L", "G"], disturbances=["X0", "Y6"],
        params=dict(Vl=5.0, Vg=5.0, m=1.0, Kla=5.0, eq_exponent=2.0, X0=0.6, Y6=0.05),
    ),
    # model_classes.py:864-931
    "four_tank": dict(
        states=["h1", "h2", "h3", "h4"], inputs=["v1", "v2"], disturbances=["None"],
        params=dict(g=9.81, gamma_1=0.2, gamma_2=0.2, k1=0.00085, k2=0.00095,
                    a1=0.0035, a2=0.0030, a3=0.0020, a4=0.0025, A1=1.0, A2=1.0, A3=1.0, A4=1.0),
    ),
    # model_classes.py:1046-1155
    "biofilm_reactor": dict(
        states=[f"{s}_{k}" for k in ("1", "2", "3", "A") for s in ("S1", "S2", "S3", "O")],
        inputs=["F", "Fr", "S1_F", "S2_F", "S3_F"], disturbances=[],
        params=dict(V=10.0, Va=15.0, Kla=1.5, m=0.5, eq_exponent=1.0, O_air=300.0, vm_1=0.8, vm_2=1.0,
                    K1=0.5, K2=0.1, KO_1=1.5, KO_2=0.5),
    ),
    # model_classes.py:1232-1345
    "crystallization": dict(
        states=["Mu0", "Mu1", "Mu2", "Mu3", "Conc", "CV", "Ln"], inputs=["Tc"],
        disturbances=["ka", "kg", "UA"],
        params=dict(ka=0.923714966, kb=-6754.878558, kc=0.92229965554, kd=1.341205945, kg=48.07514464,
                    k1=-4921.261419, k2=1.871281405, a=0.50523693, b=7.271241375, alfa=7.510905767,
                    ro=2.658),
    ),
}


def default_params(model: str) -> dict:
    return dict(MODELS[model]["params"])


def _f64(a):
    return np.asarray(a, dtype=np.float64)


# --------------------------------------------------------------------------
# RHS functions
# --------------------------------------------------------------------------
def cstr(x, u, p):
    """model_classes.py:45-62.  With 3 inputs the last two override Ti, Caf (:48-51)."""
    x, u = _f64(x), _f64(u)
    ca, T = x[0], x[1]
    Tc = u[0]
    Ti, Caf = (u[1], u[2]) if u.shape[0] == 3 else (p["Ti"], p["Caf"])
    rate = p["k0"] * np.exp(-p["EA_over_R"] / T) * ca
    thru = p["q"] / p["V"]
    dca = thru * (Caf - ca) - rate
    dT = (thru * (Ti - T)
          + ((-p["deltaHr"]) * rate) * (1.0 / (p["rho"] * p["C"]))
          + p["UA"] * (Tc - T) * (1.0 / (p["rho"] * p["C"] * p["V"])))
    return np.stack(np.broadcast_arrays(dca, dT))


def first_order_system(x, u, p):
    """model_classes.py:310-327:  dx = (K u - x) / tau."""
    x, u = _f64(x), _f64(u)
    return ((p["K"] * u[0] - x[0]) * 1.0 / p["tau"])[None]


def multistage_extraction(x, u, p):
    """model_classes.py:370-412.  Counter-current 5-stage column, states
    interleaved (X_i, Y_i).  With 4 inputs the last two override X0, Y6 (:382-385)."""
    x, u = _f64(x), _f64(u)
    L, G = u[0], u[1]
    X0, Y6 = (u[2], u[3]) if u.shape[0] == 4 else (p["X0"], p["Y6"])
    X = x[0::2]
    Y = x[1::2]
    n = X.shape[0]
    shape = np.broadcast(X[0], L, X0, Y6).shape
    out = np.empty((2 * n,) + shape)
    for i in range(n):
        xeq = (Y[i] ** p["eq_exponent"]) / p["m"]
        Q = p["Kla"] * (X[i] - xeq) * p["Vl"]
        x_in = X0 if i == 0 else X[i - 1]
        y_in = Y6 if i == n - 1 else Y[i + 1]
        out[2 * i] = (1.0 / p["Vl"]) * (L * (x_in - X[i]) - Q)
        out[2 * i + 1] = (1.0 / p["Vg"]) * (G * (y_in - Y[i]) + Q)
    return out


def four_tank(x, u, p):
    """model_classes.py:891-913.  sqrt(2 g h) outflows; NaN for h < 0 like the reference."""
    x, u = _f64(x), _f64(u)
    v1, v2 = u[0], u[1]
    with np.errstate(invalid="ignore"):
        s = np.sqrt(2.0 * p["g"] * x)
    d1 = (-p["a1"] / p["A1"]) * s[0] + (p["a3"] / p["A1"]) * s[2] + ((p["gamma_1"] * p["k1"]) / p["A1"]) * v1
    d2 = (-p["a2"] / p["A2"]) * s[1] + (p["a4"] / p["A2"]) * s[3] + ((p["gamma_2"] * p["k2"]) / p["A2"]) * v2
    d3 = (-p["a3"] / p["A3"]) * s[2] + (((1 - p["gamma_2"]) * p["k2"]) / p["A3"]) * v2
    d4 = (-p["a4"] / p["A4"]) * s[3] + (((1 - p["gamma_1"]) * p["k1"]) / p["A4"]) * v1
    return np.stack(np.broadcast_arrays(d1, d2, d3, d4))


def biofilm_reactor(x, u, p):
    """model_classes.py:1077-1136.  Three biofilm stages + absorber; Monod kinetics."""
    x, u = _f64(x), _f64(u)
    F, Fr, feed = u[0], u[1], u[2:5]
    stage = x.reshape((4, 4) + x.shape[1:])          # [unit(1,2,3,A)][S1,S2,S3,O]
    shape = np.broadcast(x[0], F).shape
    out = np.empty((16,) + shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(3):
            S1, S2, S3, O = stage[k]
            r1 = ((p["vm_1"] * S1) / (p["K1"] + S1)) * (O / (p["KO_1"] + O))
            r2 = ((p["vm_2"] * S2) / (p["K2"] + S2)) * (O / (p["KO_2"] + O))
            ro = -r1 * 3.5 - r2 * 1.1
            upstream = stage[3] if k == 0 else stage[k - 1]
            gen = (-r1, r1 - r2, r2, ro)
            for j in range(4):
                out[4 * k + j] = (Fr / p["V"]) * (upstream[j] - stage[k][j]) - gen[j]
    O_eq = (p["O_air"] ** p["eq_exponent"]) / p["m"]
    for j in range(3):
        out[12 + j] = (Fr / p["Va"]) * (stage[2][j] - stage[3][j]) + (F / p["Va"]) * (feed[j] - stage[3][j])
    out[15] = (Fr / p["Va"]) * (stage[2][3] - stage[3][3]) + p["Kla"] * (O_eq - stage[3][3])
    return out


def crystallization(x, u, p):
    """model_classes.py:1272-1319.  Moment model; only u[0] (temperature, degC) is read (:1299)."""
    x, u = _f64(x), _f64(u)
    mu0, mu1, mu2, mu3, conc = x[0], x[1], x[2], x[3], x[4]
    TK = u[0] + 273.15
    Ceq = -686.2686 + 3.579165 * TK - 0.00292874 * TK ** 2
    S = conc * 1e3 - Ceq
    with np.errstate(invalid="ignore", divide="ignore"):
        B0 = p["ka"] * np.exp(p["kb"] / TK) * (S ** 2) ** (p["kc"] / 2) * ((mu3 ** 2) ** (p["kd"] / 2))
        Ginf = p["kg"] * np.exp(p["k1"] / TK) * (S ** 2) ** (p["k2"] / 2)
        d0 = B0
        d1 = Ginf * (p["a"] * mu0 + p["b"] * mu1 * 1e-4) * 1e4
        d2 = 2 * Ginf * (p["a"] * mu1 * 1e-4 + p["b"] * mu2 * 1e-8) * 1e8
        d3 = 3 * Ginf * (p["a"] * mu2 * 1e-8 + p["b"] * mu3 * 1e-12) * 1e12
        dc = -0.5 * p["ro"] * p["alfa"] * Ginf * (p["a"] * mu2 * 1e-8 + p["b"] * mu3 * 1e-12)
        CV = np.sqrt(mu2 * mu0 / (mu1 ** 2) - 1)
        dCV = 1 / (2 * CV + 1e-10) * ((d2 * mu0 + mu2 * d0) * mu1 ** 2 - mu2 * mu0 * 2 * mu1 * d1) / (mu1 ** 4 + 1e-10)
        dLn = (d1 * mu0 - mu1 * d0) / (mu0 ** 2 + 1e-10)
    return np.stack(np.broadcast_arrays(d0, d1, d2, d3, dc, dCV, dLn))


RHS = {
    "cstr": cstr,
    "first_order_system": first_order_system,
    "multistage_extraction": multistage_extraction,
    "four_tank": four_tank,
    "biofilm_reactor": biofilm_reactor,
    "crystallization": crystallization,
}


def rhs(model: str, x, u, p=None):
    """Evaluate model `model` at (x, u); `p` defaults to the reference defaults."""
    if p is None:
        p = MODELS[model]["params"]
    return RHS[model](x, u, p)
