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
