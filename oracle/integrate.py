"""Oracle integrators: one zero-order-hold control interval x(t) -> x(t+dt).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates `/root/reference/src/pcgym/integrator.py:90-107,163-182` (casadi_step:
autonomous RHS, input `p=uk` held constant over [0, dt], returns `xf`) and
`:65-88` (jax_step: same map via an explicit RK5(4) pair).  CasADi/SUNDIALS is
not installable in this image, so the map is evaluated with converged scipy
integrators instead; DESIGN.md explains why a *converged* solution is the
parity target (CVODES at its default tolerances is itself only ~1e-5 accurate).

  advance_truth   scipy DOP853, rtol 1e-12 / atol 1e-14, one env at a time.
  advance_dense   classical RK4 with many substeps, vectorised over envs
                  (used for batch-sized parity; error ~ (dt/n)^4).
  advance_lsoda   scipy LSODA rtol 1e-6 / atol 1e-8: stand-in for the
                  reference's CVODES defaults, used only for CPU timing.
  advance_fixed   n_sub steps of a NAMED explicit pair without step control
                  ("rk4", "dopri5", "tsit5"), vectorised.  The reference's jax
                  path is diffrax Tsit5 (integrator.py:57); a kernel that takes
                  one Tsit5 step per control interval is compared with this at
                  ~1e-12 (same scheme, different arithmetic order) and with the
                  converged solution at the north-star rtol 1e-6.
  advance_adaptive the embedded pairs ("dopri5", "tsit5") WITH the kernels' step-size
                  controller (Hairer/Norsett/Wanner II.4 integral controller, FSAL, step
                  carried between intervals), one env at a time: the same-scheme check
                  of make_env's DEFAULT integrator -- the stand-in for the reference's
                  jax path, diffrax Tsit5 + PIDController(rtol 1e-8, atol 1e-8)
                  (integrator.py:57-88), which is an integral controller as well
                  (pcoeff = dcoeff = 0) -- including its count of right-hand sides.
  advance_rodas4  the RODAS4 Rosenbrock 4(3) method with step-size control and a
                  finite-difference Jacobian, one env at a time: the same-scheme
                  check of the kernels' stiff integrator (scheme "rodas4"), which
                  stands in for the reference's CVODES BDF (integrator.py:163-182)
                  where the dynamics are stiff (an ignited cstr).  The converged
                  reference for such cases is advance_truth(method="Radau").
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from . import rhs as _rhs


def _as_fun(model, p):
    """`model` is a registry name or a reference-style callable model(x, u)."""
    if callable(model):
        return lambda x, u: np.asarray(model(x, u), dtype=np.float64)
    if p is None:
        p = _rhs.default_params(model)
    f = _rhs.RHS[model]
    return lambda x, u: f(x, u, p)


def advance_truth(model, x, u, dt, p=None, rtol=1e-12, atol=1e-14, method="DOP853"):
    """Converged x(dt) for ONE env: x[nx], u[nu].  integrator.py:102-106 semantics."""
    f = _as_fun(model, p)
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    sol = solve_ivp(lambda t, y: f(y, u), (0.0, float(dt)), x, method=method, rtol=rtol, atol=atol)
    if not sol.success:
        raise RuntimeError(f"oracle integrator failed: {sol.message}")
    return sol.y[:, -1]


def advance_lsoda(model, x, u, dt, p=None, rtol=1e-6, atol=1e-8):
    """CVODES-default stand-in (CasADi SundialsInterface defaults reltol 1e-6, abstol 1e-8)."""
    return advance_truth(model, x, u, dt, p, rtol=rtol, atol=atol, method="LSODA")


def advance_dense(model, x, u, dt, p=None, n_sub=64):
    """Vectorised RK4 with `n_sub` substeps: x[nx, N], u[nu, N] -> x[nx, N]."""
    f = _as_fun(model, p)
    x = np.array(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    h = float(dt) / n_sub
    for _ in range(n_sub):
        k1 = f(x, u)
        k2 = f(x + (0.5 * h) * k1, u)
        k3 = f(x + (0.5 * h) * k2, u)
        k4 = f(x + h * k3, u)
        x = x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return x


# Butcher tableaus, restated from the literature (NOT parsed from the CUDA sources: tests/test_tableaus.py checks
# both against the order conditions and against each other).  Rows 1..S-1 are the stage rows, B the solution weights.
#   rk4    : Kutta 1901.
#   dopri5 : Dormand & Prince 1980, RK5(4)7M (scipy's "RK45").
#   tsit5  : Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the first column simplifying
#            assumption" -- the pair diffrax.Tsit5 implements (reference integrator.py:57).
TABLEAUS = {
    "rk4": {
        "A": [[], [1 / 2], [0, 1 / 2], [0, 0, 1]],
        "B": [1 / 6, 1 / 3, 1 / 3, 1 / 6],
    },
    "dopri5": {
        "A": [[], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
              [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
              [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]],
        "B": [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
    },
    "tsit5": {
        "A": [[], [0.161],
              [-0.008480655492356989, 0.335480655492357],
              [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
              [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
              [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]],
        "B": [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
              2.324710524099774],
    },
}


def advance_fixed(model, x, u, dt, p=None, scheme="tsit5", n_sub=1):
    """`n_sub` fixed steps of the explicit scheme, input held: x[nx, N], u[nu, N] -> x[nx, N]."""
    f = _as_fun(model, p)
    tb = TABLEAUS[scheme]
    x = np.array(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    h = float(dt) / n_sub
    for _ in range(n_sub):
        k = []
        for row in tb["A"]:
            xi = x.copy()
            for a, kj in zip(row, k):
                if a != 0.0:
                    xi = xi + (h * a) * kj
            k.append(f(xi, u))
        for b, kj in zip(tb["B"], k):
            if b != 0.0:
                x = x + (h * b) * kj
    return x


# error weights b - b_hat of the embedded fourth-order solutions, over the S stages and the FSAL stage f(x_new)
TABLEAUS["dopri5"]["E"] = [71 / 57600, 0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40]
TABLEAUS["tsit5"]["E"] = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                          0.5823571654525552, -0.45808210592918697, 1 / 66]


def advance_adaptive(model, x, u, dt, p=None, scheme="tsit5", rtol=1e-8, atol=1e-8, h0=None, max_steps=4096):
    """x(dt) for ONE env by the embedded pair with the kernels' controller (csrc/rollout_kernel.cuh:
    advance_adaptive_aux): err^2 = mean((e / (atol + rtol max(|x|, |x'|)))^2), accept if err <= 1,
    h *= clamp(0.9 err^-1/5, 0.2, 10) (no growth right after a rejection), a truncated last step does not shrink the
    carried step, the derivative at x_new is reused as the next first stage.  Returns (x, h_carry, n_rhs)."""
    f = _as_fun(model, p)
    tb = TABLEAUS[scheme]
    x = np.asarray(x, dtype=np.float64).copy()
    u = np.asarray(u, dtype=np.float64)
    t, h = 0.0, (h0 if (h0 is not None and 0 < h0 < dt) else float(dt))
    k0, n_rhs, steps, rejected = f(x, u), 1, 0, False
    while t < dt:
        steps += 1
        if steps > max_steps:
            raise RuntimeError("advance_adaptive: step cap")
        left = dt - t
        last = h >= left * (1 - 1e-12)
        hh = left if last else h
        k = [k0]
        for row in tb["A"][1:]:
            xi = x.copy()
            for a, kj in zip(row, k):
                if a != 0.0:
                    xi = xi + (hh * a) * kj
            k.append(f(xi, u))
        xn = x.copy()
        for b, kj in zip(tb["B"], k):
            if b != 0.0:
                xn = xn + (hh * b) * kj
        kn = f(xn, u)
        n_rhs += len(tb["B"])
        err = sum(e * (hh * kj) for e, kj in zip(tb["E"], k + [kn]) if e != 0.0)
        sc = atol + rtol * np.maximum(np.abs(x), np.abs(xn))
        msq = float(np.mean((err / sc) ** 2))
        grow = 0.9 * max(msq, 1e-30) ** -0.1 if np.isfinite(msq) else 0.2
        if np.isfinite(msq) and msq <= 1.0:
            t = dt if last else t + hh
            x, k0 = xn, kn
            hn = hh * min(1.0 if rejected else 10.0, grow)
            h = max(h, hn) if last else hn
            rejected = False
        else:
            h = hh * max(0.2, min(grow, 1.0))
            rejected = True
            if not h > 1e-14 * dt:
                raise RuntimeError("advance_adaptive: step size underflow")
    return x, h, n_rhs


# RODAS4 (Hairer & Wanner, Solving ODEs II, sec. IV.7; coefficients of rodas.f), transformed form:
#   W = I / (gamma h) - J(y0);   W U_i = f(y0 + sum_j a_ij U_j) + sum_j (c_ij / h) U_j;   y1 = y0 + sum_j a_5j U_j + U_6
RODAS4 = {
    "gamma": 0.25,
    "A": [[],
          [0.1544e1],
          [0.9466785280815826, 0.2557011698983284],
          [0.3314825187068521e1, 0.2896124015972201e1, 0.9986419139977817],
          [0.1221224509226641e1, 0.6019134481288629e1, 0.1253708332932087e2, -0.6878860361058950],
          [0.1221224509226641e1, 0.6019134481288629e1, 0.1253708332932087e2, -0.6878860361058950, 1.0]],
    "C": [[],
          [-0.56688e1],
          [-0.2430093356833875e1, -0.2063599157091915],
          [-0.1073529058151375, -0.9594562251023355e1, -0.2047028614809616e2],
          [0.7496443313967647e1, -0.1024680431464352e2, -0.3399990352819905e2, 0.1170890893206160e2],
          [0.8083246795921522e1, -0.7981132988064893e1, -0.3152159432874371e2, 0.1631930543123136e2,
           -0.6058818238834054e1]],
}


def _fd_jacobian(f, x, u):
    """Central differences with a state-scaled increment (~1e-9 relative: the effect on a RODAS4 step is far below
    the step-size controller's tolerance; the kernels use the analytic Jacobian)."""
    n = x.shape[0]
    J = np.empty((n, n))
    for j in range(n):
        d = 1e-6 * max(abs(x[j]), 1e-3)
        e = np.zeros(n)
        e[j] = d
        J[:, j] = (f(x + e, u) - f(x - e, u)) / (2 * d)
    return J


def rodas4_step(f, J, x, u, h):
    """One RODAS4 step of size h from x with Jacobian matrix J: returns (x_new, error estimate U_6)."""
    n = x.shape[0]
    W = np.eye(n) / (RODAS4["gamma"] * h) - J
    U = []
    xi = x
    for a_row, c_row in zip(RODAS4["A"], RODAS4["C"]):
        xi = x + sum((a * Uj for a, Uj in zip(a_row, U)), np.zeros(n))
        rhs = f(xi, u) + sum(((c / h) * Uj for c, Uj in zip(c_row, U)), np.zeros(n))
        U.append(np.linalg.solve(W, rhs))
    return xi + U[-1], U[-1]


def advance_rodas4(model, x, u, dt, p=None, rtol=1e-8, atol=1e-8, h0=None, jac=None, max_steps=100000):
    """x(dt) for ONE env by step-size-controlled RODAS4 (controller of the kernels: err = RMS(U_6 / (atol + rtol max(|x|,
    |x'|))), h *= clamp(0.9 err^-1/4, 0.2, 6), no growth right after a rejection).  Returns (x, h_carry, steps)."""
    f = _as_fun(model, p)
    x = np.asarray(x, dtype=np.float64).copy()
    u = np.asarray(u, dtype=np.float64)
    t, h = 0.0, (h0 if (h0 is not None and 0 < h0 < dt) else 0.25 * dt)
    steps, rejected = 0, False
    while t < dt:
        steps += 1
        if steps > max_steps:
            raise RuntimeError("advance_rodas4: step cap")
        left = dt - t
        last = h >= left * (1 - 1e-12)
        hh = left if last else h
        J = jac(x, u) if jac is not None else _fd_jacobian(f, x, u)
        xn, err = rodas4_step(f, J, x, u, hh)
        sc = atol + rtol * np.maximum(np.abs(x), np.abs(xn))
        msq = float(np.mean((err / sc) ** 2))
        grow = 0.9 * max(msq, 1e-30) ** -0.125 if np.isfinite(msq) else 0.2
        if np.isfinite(msq) and msq <= 1.0:
            t = dt if last else t + hh
            x = xn
            hn = hh * min(1.0 if rejected else 6.0, grow)
            h = max(h, hn) if last else hn
            rejected = False
        else:
            h = hh * max(0.2, min(grow, 1.0))
            rejected = True
    return x, h, steps
