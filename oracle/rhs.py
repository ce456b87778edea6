"""Oracle right-hand sides: numpy restatement of the registry models (the six north-star ones first).

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
    # extension: ten stages of the same balances (see multistage_extraction_10stage)
    "multistage_extraction_10stage": dict(
        states=[f"{s}{k}" for k in range(1, 11) for s in ("X", "Y")],
        inputs=["L", "G"], disturbances=["X0", "Y11"],
        params=dict(Vl=5.0, Vg=5.0, m=1.0, Kla=5.0, eq_exponent=2.0, X0=0.6, Y11=0.05),
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


def multistage_extraction_10stage(x, u, p):
    """EXTENSION (BASELINE.json config 4 words the workload "10-stage, ~20-state"; the reference column has five
    stages, model_classes.py:381,424): the SAME stage balances (:387-410) for ten stages, gas feed Y11.  There is no
    reference implementation of this size; tests pin it through the five-stage code path it shares and the column's
    overall mass balance."""
    q = dict(p)
    q["Y6"] = q.pop("Y11")
    return multistage_extraction(x, u, q)


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



# --------------------------------------------------------------------------
# further registry models; keys are the registry names of pcgym.py:128-148
# --------------------------------------------------------------------------
MODELS.update({
    # model_classes.py:65-125
    "complex_cstr": dict(
        states=["Ca", "Cb", "Cc", "T"], inputs=["Tc"], disturbances=["Ti", "Caf"],
        params=dict(q=100.0, V=100.0, rho=1000.0, C=0.239, deltaHr1=-5e4, EA1_over_R=8750.0, k01=7.2e10,
                    deltaHr2=-3e4, EA2_over_R=9000.0, k02=1.0e10, UA=5e4, Ti=350.0, Caf=1.0)),
    # model_classes.py:509-558
    "nonsmooth_control": dict(
        states=["X1", "X2"], inputs=["U"], disturbances=["None"],
        params=dict(a_11=0.0, a_12=1.0, a_21=-2.0, a_22=-3.0, b_1=0.0, b_2=1.0)),
    # model_classes.py:611-679
    "cstr_series_recycle": dict(
        states=["C1", "T1", "C2", "T2"], inputs=["F", "L", "Tc1", "Tc2"], disturbances=[],
        params=dict(C_O=97.35, T_O=298.0, V1=1e-3, V2=2e-3, U1A1=0.461, U2A2=0.732, rho=1.05e3, cp=3.766,
                    k=3.118e5, E=46.14, deltaH=58.41, R=8.3145e-3)),
    # model_classes.py:682-760
    "distillation_column": dict(
        states=["X0", "X1", "X2", "X3", "Xf", "X4", "X5", "X6", "Xb"], inputs=["R", "F"], disturbances=[],
        params=dict(D=100.0, q=1.0, alpha=5.0, X_feed=0.2, M0=2000.0, Mb=2000.0, M=2000.0)),
    # model_classes.py:763-861
    "multistage_extraction_reactive": dict(
        states=[f"{s}{k}" for k in range(1, 6) for s in ("XA", "YA", "YB", "YC")], inputs=["L", "G"],
        disturbances=[],
        params=dict(Vl=5.0, Vg=5.0, m=1.0, Kla=0.01, k=0.1, eq_exponent=2.0, XA0=2.0, YA6=0.0, YB6=2.0, YC6=0.0)),
    # model_classes.py:935-1043 (info() has no "disturbances" key)
    "heat_exchanger": dict(
        states=[f"{s}{k}" for k in range(1, 9) for s in ("Tt", "Tm", "Ts")], inputs=["Ft", "Fs", "Tt0", "Ts9"],
        disturbances=[],
        params=dict(Utm=1.0, Usm=1.0, L=1.0, Dt=1.0, Dm=2.0, Ds=3.0, cpt=1.0, cpm=1.0, cps=1.0, rhot=1.0,
                    rhom=1.0, rhos=1.0)),
    # model_classes.py:1158-1229
    "polymerisation_reactor": dict(
        states=["T", "M", "I"], inputs=["F", "Tf", "Mf", "If"], disturbances=[],
        params=dict(Ap=6e10, Ad=4e10, At=9e10, Ep_over_R=7750.0, Ed_over_R=8500.0, Et_over_R=8250.0, f=0.5, V=1.0,
                    deltaHp=-3e4, rho=1200.0, cp=2.0)),
    # model_classes.py:433-506 (class photo_production)
    "photobioreactor": dict(
        states=["c_x", "c_N", "c_q"], inputs=["I", "F_N"], disturbances=[],
        params=dict(u_m=0.0572, u_d=0.0, Y_NX=504.5, k_m=0.00016, k_d=0.281, k_sq=23.51, K_Nq=16.89, k_iq=800.0,
                    k_s=178.9, k_i=447.1, k_N=393.1)),
    # model_classes.py:268-293
    "invariant_batch": dict(
        states=["xA", "xB", "xC", "xD"], inputs=[], disturbances=[],
        params=dict(k1f=55.0, k1r=1.0, k2f=2.0, k2r=1.0)),
    # model_classes.py:222-263
    "batch": dict(
        states=["Ca", "Cb", "Cc", "T"], inputs=["Tc"], disturbances=[],
        params=dict(k01=1.0, k02=0.5, EA1=5000.0, EA2=6000.0, R=8.314, dH1=-1000.0, dH2=-1500.0, rho=1000.0,
                    Cp=4.0, UA=100.0, V=1.0)),
    # model_classes.py:186-216 (class coupled_oscillators)
    "coupled_oscillator": dict(
        states=[f"x{i + 1}" for i in range(10)] + [f"p{i + 1}" for i in range(10)], inputs=[], disturbances=[],
        params=dict(N=10, k=1.0, m=1.0)),
    # model_classes.py:156-180 (class disease_model)
    "disease": dict(states=["S", "I", "R"], inputs=["u"], disturbances=[], params=dict(beta=0.3, gamma=0.1)),
    # model_classes.py:128-150
    "hydraulic_tank": dict(states=["q1", "q2"], inputs=["u"], disturbances=[], params=dict(D=1.0)),
})


def _stack(rows):
    return np.stack(np.broadcast_arrays(*rows))


def complex_cstr(x, u, p):
    """model_classes.py:102-125.  With 3 inputs the last two override Ti, Caf (:105-108)."""
    x, u = _f64(x), _f64(u)
    ca, cb, cc, T = x[0], x[1], x[2], x[3]
    Tc = u[0]
    Ti, Caf = (u[1], u[2]) if u.shape[0] == 3 else (p["Ti"], p["Caf"])
    r1 = p["k01"] * np.exp(-p["EA1_over_R"] / T) * ca
    r2 = p["k02"] * np.exp(-p["EA2_over_R"] / T) * cb
    thru = p["q"] / p["V"]
    heat_gen = (-p["deltaHr1"] * r1) + (-p["deltaHr2"] * r2)
    return _stack([thru * (Caf - ca) - r1, thru * (0 - cb) + 2 * r1 - r2, thru * (0 - cc) + r2,
                   thru * (Ti - T) + heat_gen / (p["rho"] * p["C"]) + (p["UA"] / (p["rho"] * p["C"] * p["V"])) * (Tc - T)])


def nonsmooth_control(x, u, p):
    """model_classes.py:535-541:  linear 2-state system, scalar input."""
    x, u = _f64(x), _f64(u)
    return _stack([p["a_11"] * x[0] + p["a_12"] * x[1] + p["b_1"] * u[0],
                   p["a_21"] * x[0] + p["a_22"] * x[1] + p["b_2"] * u[0]])


def cstr_series_recycle(x, u, p):
    """model_classes.py:646-655."""
    x, u = _f64(x), _f64(u)
    C1, T1, C2, T2 = x[0], x[1], x[2], x[3]
    F, L, Tc1, Tc2 = u[0], u[1], u[2], u[3]
    e1, e2 = np.exp(-p["E"] / (p["R"] * T1)), np.exp(-p["E"] / (p["R"] * T2))
    rc = p["rho"] * p["cp"]
    heat = (p["k"] * (-p["deltaH"])) / rc
    return _stack([
        (p["C_O"] / p["V1"]) * F + (1 / p["V1"]) * L * C2 - (1 / p["V1"]) * (F + L) * C1 - p["k"] * C1 * e1,
        (p["T_O"] / p["V1"]) * F + (1 / p["V1"]) * L * T2 - (p["U1A1"] / (p["V1"] * rc)) * (T1 - Tc1)
        - (1 / p["V1"]) * (F + L) * T1 + heat * C1 * e1,
        (1 / p["V2"]) * (F + L) * (C1 - C2) - p["k"] * C2 * e2,
        (1 / p["V2"]) * (F + L) * (T1 - T2) - (p["U2A2"] / (p["V2"] * rc)) * (T2 - Tc2) + heat * C2 * e2])


def distillation_column(x, u, p):
    """model_classes.py:706-734.  Rectifying / feed / stripping trays with constant relative volatility."""
    x, u = _f64(x), _f64(u)
    R, F = u[0], u[1]
    L = R * p["D"]
    V = (R + 1) * p["D"]
    Ld = L + p["q"] * F
    Vd = V + (1 - p["q"]) * F
    W = F - p["D"]
    Y = [(p["alpha"] * x[i]) / (1 + (p["alpha"] - 1) * x[i]) for i in range(9)]
    M = p["M"]
    return _stack([
        (1 / p["M0"]) * ((V * Y[1]) - (L + p["D"]) * x[0]),
        (1 / M) * (L * (x[0] - x[1]) + V * (Y[2] - Y[1])),
        (1 / M) * (L * (x[1] - x[2]) + V * (Y[3] - Y[2])),
        (1 / M) * (L * (x[2] - x[3]) + V * (Y[4] - Y[3])),
        (1 / M) * (L * x[3] - Ld * x[4] + Vd * Y[5] - V * Y[4] + F * p["X_feed"]),
        (1 / M) * (Ld * (x[4] - x[5]) + Vd * (Y[6] - Y[5])),
        (1 / M) * (Ld * (x[5] - x[6]) + Vd * (Y[7] - Y[6])),
        (1 / M) * (Ld * (x[6] - x[7]) + Vd * (Y[8] - Y[7])),
        (1 / p["Mb"]) * (Ld * x[7] - W * x[8] - Vd * Y[8])])


def multistage_extraction_reactive(x, u, p):
    """model_classes.py:793-841.  5 stages x (XA, YA, YB, YC); reaction A + B -> C in the gas phase."""
    x, u = _f64(x), _f64(u)
    L, G = u[0], u[1]
    rows = []
    feed_g = (p["YA6"], p["YB6"], p["YC6"])
    for i in range(5):
        XA, YA, YB, YC = x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]
        Q = p["Kla"] * (XA - ((YA ** p["eq_exponent"]) / p["m"])) * p["Vl"]
        r = p["k"] * YA * YB
        xin = p["XA0"] if i == 0 else x[4 * i - 4]
        ya, yb, yc = feed_g if i == 4 else (x[4 * i + 5], x[4 * i + 6], x[4 * i + 7])
        rows += [(1 / p["Vl"]) * (L * (xin - XA) - Q),
                 (1 / p["Vg"]) * (G * (ya - YA) + Q - r * p["Vg"]),
                 (1 / p["Vg"]) * (G * (yb - YB) - r * p["Vg"]),
                 (1 / p["Vg"]) * (G * (yc - YC) + r * p["Vg"])]
    return _stack(rows)


def heat_exchanger(x, u, p):
    """model_classes.py:966-1017.  8 cells x (tube, metal, shell); counter-current."""
    x, u = _f64(x), _f64(u)
    Ft, Fs, Tt0, Ts9 = u[0], u[1], u[2], u[3]
    Vt = p["L"] * np.pi * p["Dt"] ** 2
    At = p["L"] * np.pi * p["Dt"]
    Vm = p["L"] * np.pi * (p["Dm"] ** 2 - p["Dt"] ** 2)
    Am = p["L"] * np.pi * p["Dm"]
    Vs = p["L"] * np.pi * (p["Ds"] ** 2 - p["Dm"] ** 2)
    rows = []
    for i in range(8):
        Tt, Tm, Ts = x[3 * i], x[3 * i + 1], x[3 * i + 2]
        Qt = p["Utm"] * At * (Tt - Tm)
        Qm = p["Usm"] * Am * (Tm - Ts)
        tt_in = Tt0 if i == 0 else x[3 * i - 3]
        ts_in = Ts9 if i == 7 else x[3 * i + 5]
        rows += [(1 / (p["cpt"] * p["rhot"] * Vt)) * (Ft * p["cpt"] * (tt_in - Tt) - Qt),
                 (1 / (p["cpm"] * p["rhom"] * Vm)) * (Qt - Qm),
                 (1 / (p["cps"] * p["rhos"] * Vs)) * (Fs * p["cps"] * (ts_in - Ts) + Qm)]
    return _stack(rows)


def polymerisation_reactor(x, u, p):
    """model_classes.py:1196-1211."""
    x, u = _f64(x), _f64(u)
    T, M, I = x[0], x[1], x[2]
    F, Tf, Mf, If = u[0], u[1], u[2], u[3]
    kp = p["Ap"] * np.exp(-p["Ep_over_R"] / T)
    kd = p["Ad"] * np.exp(-p["Ed_over_R"] / T)
    kt = p["At"] * np.exp(-p["Et_over_R"] / T)
    ri = 2 * p["f"] * kd * I
    with np.errstate(invalid="ignore"):
        rp = kp * ((p["f"] * kd * I) / kt) ** 0.5
    return _stack([(F / p["V"]) * (Tf - T) + ((-p["deltaHp"]) / (p["rho"] * p["cp"])) * rp,
                   (F / p["V"]) * (Mf - M) - rp,
                   (F / p["V"]) * (If - I) - ri])


def photobioreactor(x, u, p):
    """model_classes.py:477-486 (class photo_production)."""
    x, u = _f64(x), _f64(u)
    c_x, c_N, c_q = x[0], x[1], x[2]
    I, F_N = u[0], u[1]
    den = I + p["k_s"] + (I ** 2 / p["k_i"])                   # evaluation order as the reference (left to right)
    return _stack([p["u_m"] * I / den * c_x * c_N / (c_N + p["k_N"]) - p["u_d"] * c_x,
                   -p["Y_NX"] * p["u_m"] * I / den * c_x * c_N / (c_N + p["k_N"]) + F_N,
                   p["k_m"] * I / (I + p["k_sq"] + (I ** 2 / p["k_iq"])) * c_x - (p["k_d"] * c_q) / (c_N + p["K_Nq"])])


def invariant_batch(x, u, p):
    """model_classes.py:284-291 (no inputs)."""
    x = _f64(x)
    xA, xB, xC, xD = x[0], x[1], x[2], x[3]
    q1 = p["k1f"] * xA * xB - p["k1r"] * xC
    q2 = p["k2f"] * xA * xC - p["k2r"] * xD
    return _stack([-q1 - q2, -q1, q1 - q2, q2])


def batch(x, u, p):
    """model_classes.py:245-261."""
    x, u = _f64(x), _f64(u)
    CA, CB, CC, T = x[0], x[1], x[2], x[3]
    Tc = u[0]
    r1 = p["k01"] * np.exp(-p["EA1"] / (p["R"] * T)) * CA
    r2 = p["k02"] * np.exp(-p["EA2"] / (p["R"] * T)) * CB
    return _stack([-r1, 2 * r1 - r2, r2,
                   -(p["dH1"] * r1 + p["dH2"] * r2) / (p["rho"] * p["Cp"])
                   + p["UA"] / (p["rho"] * p["Cp"] * p["V"]) * (Tc - T)])


def coupled_oscillator(x, u, p):
    """model_classes.py:202-216: ring of N masses, positions then momenta (no inputs)."""
    x = _f64(x)
    N = int(p["N"])
    pos, mom = x[:N], x[N:]
    rows = [mom[i] / p["m"] for i in range(N)]
    rows += [-p["k"] * (2 * pos[i] - pos[(i - 1) % N] - pos[(i + 1) % N]) for i in range(N)]
    return _stack(rows)


def disease(x, u, p):
    """model_classes.py:170-178 (class disease_model): SIRS with vaccination."""
    x, u = _f64(x), _f64(u)
    S, I = x[0], x[1]
    return _stack([-p["beta"] * S * I - u[0] * S, p["beta"] * S * I - p["gamma"] * I, p["gamma"] * I + u[0] * S])


def hydraulic_tank(x, u, p):
    """model_classes.py:141-148."""
    x, u = _f64(x), _f64(u)
    return _stack([-p["D"] * (x[0] - x[1]) + u[0], p["D"] * (x[0] - x[1]) - u[0]])


RHS = {
    "cstr": cstr,
    "first_order_system": first_order_system,
    "multistage_extraction": multistage_extraction,
    "multistage_extraction_10stage": multistage_extraction_10stage,
    "four_tank": four_tank,
    "biofilm_reactor": biofilm_reactor,
    "crystallization": crystallization,
    "complex_cstr": complex_cstr,
    "nonsmooth_control": nonsmooth_control,
    "cstr_series_recycle": cstr_series_recycle,
    "distillation_column": distillation_column,
    "multistage_extraction_reactive": multistage_extraction_reactive,
    "heat_exchanger": heat_exchanger,
    "polymerisation_reactor": polymerisation_reactor,
    "photobioreactor": photobioreactor,
    "invariant_batch": invariant_batch,
    "batch": batch,
    "coupled_oscillator": coupled_oscillator,
    "disease": disease,
    "hydraulic_tank": hydraulic_tank,
}


def rhs(model: str, x, u, p=None):
    """Evaluate model `model` at (x, u); `p` defaults to the reference defaults."""
    if p is None:
        p = MODELS[model]["params"]
    return RHS[model](x, u, p)
