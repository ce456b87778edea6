"""Model registry: host-side mirror of the reference's `model_mapping` (src/pcgym/pcgym.py:128-148)
and of each model's `info()` (src/pcgym/model_classes.py).

The arithmetic lives in the CUDA library (pcgym_b200/csrc/models.cuh); these objects only carry the
names, ordering and parameter values that make_env needs, with the reference's duck-typed surface:
attribute access to parameters (`model.k0`), `info()`, `int_method`.
"""
from __future__ import annotations

from . import _native

# names in reference order: states, manipulated inputs, disturbance channels
_META = {
    "cstr": (["Ca", "T"], ["Tc"], ["Ti", "Caf"]),                                         # :23-43
    "first_order_system": (["x"], ["u"], ["None"]),                                       # :329-343
    "multistage_extraction": (["X1", "Y1", "X2", "Y2", "X3", "Y3", "X4", "Y4", "X5", "Y5"],
                              ["L", "G"], ["X0", "Y6"]),                                   # :415-430
    "four_tank": (["h1", "h2", "h3", "h4"], ["v1", "v2"], ["None"]),                      # :915-931
    "crystallization": (["Mu0", "Mu1", "Mu2", "Mu3", "Conc", "CV", "Ln"], ["Tc"], ["ka", "kg", "UA"]),   # :1323-1345
    "biofilm_reactor": ([f"{s}_{k}" for k in ("1", "2", "3", "A") for s in ("S1", "S2", "S3", "O")],
                        ["F", "Fr", "S1_F", "S2_F", "S3_F"], []),                          # :1138-1155
}

# ---- further registry models (csrc/models_ext.cuh).  `info()["parameters"]` holds the numeric model parameters only
# (the reference copies `self.__dict__`, which for some classes also carries "int_method" or the name lists);
# kept from the reference: heat_exchanger has no "disturbances" key, photo_production adds "uncertainties": [].
_META.update({
    "complex_cstr": (["Ca", "Cb", "Cc", "T"], ["Tc"], ["Ti", "Caf"]),                        # :97-100
    "nonsmooth_control": (["X1", "X2"], ["U"], ["None"]),                                   # :551-556
    "cstr_series_recycle": (["C1", "T1", "C2", "T2"], ["F", "L", "Tc1", "Tc2"], []),        # :671-677
    "distillation_column": (["X0", "X1", "X2", "X3", "Xf", "X4", "X5", "X6", "Xb"], ["R", "F"], []),   # :752-758
    "multistage_extraction_reactive": ([f"{s}{k}" for k in range(1, 6) for s in ("XA", "YA", "YB", "YC")],
                                       ["L", "G"], []),                                     # :853-859
    "heat_exchanger": ([f"{s}{k}" for k in range(1, 9) for s in ("Tt", "Tm", "Ts")],
                       ["Ft", "Fs", "Tt0", "Ts9"], None),                                   # :1036-1041 (no key)
    "polymerisation_reactor": (["T", "M", "I"], ["F", "Tf", "Mf", "If"], []),               # :1222-1228
    "photobioreactor": (["c_x", "c_N", "c_q"], ["I", "F_N"], []),                           # :497-505
    "invariant_batch": (["xA", "xB", "xC", "xD"], [], []),                                  # :279-282
    "batch": (["Ca", "Cb", "Cc", "T"], ["Tc"], []),                                         # :240-243
    "coupled_oscillator": ([f"x{i + 1}" for i in range(10)] + [f"p{i + 1}" for i in range(10)], [], []),   # :197-200
    "disease": (["S", "I", "R"], ["u"], []),                                                # :166-169
    "hydraulic_tank": (["q1", "q2"], ["u"], []),                                            # :136-139
    # EXTENSION (not in the reference's registry): ten stages of multistage_extraction's balances, gas feed Y11
    "multistage_extraction_10stage": ([f"{s}{k}" for k in range(1, 11) for s in ("X", "Y")], ["L", "G"], ["X0", "Y11"]),
})
_NOT_YET = []


class DeviceModel:
    """A registry model whose RHS runs in the CUDA library."""

    def __init__(self, name: str, int_method: str = "b200", **overrides):
        if name not in _META:
            raise KeyError(name)
        self.name = name
        self.model_id = _native.MODEL_IDS[name]
        self.int_method = int_method
        self.states, self.inputs, self.disturbances = (None if v is None else list(v) for v in _META[name])
        params = _native.model_params(self.model_id)
        for k, v in overrides.items():
            if k not in params:
                raise AttributeError(f"{name} has no parameter {k!r}")
            params[k] = float(v)
        self.__dict__["_params"] = params
        nx, nu, nd, _ = _native.model_dims(self.model_id)
        real_nd = 0 if self.disturbances in (["None"], [], None) else len(self.disturbances)
        # input-less models carry one ignored input row on the device (csrc/models_ext.cuh)
        assert (nx, nu, nd) == (len(self.states), max(len(self.inputs), 1), real_nd), "registry / library mismatch"
        if name == "coupled_oscillator" and int(params["N"]) != 10:
            raise NotImplementedError("coupled_oscillator: the state count is 2 N and only N = 10 has a device kernel")

    def __getattr__(self, k):
        p = self.__dict__.get("_params", {})
        if k in p:
            return p[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        p = self.__dict__.get("_params")
        if p is not None and k in p:
            p[k] = float(v)
        else:
            self.__dict__[k] = v

    def theta(self):
        return list(self._params.values())

    def info(self) -> dict:
        out = {"parameters": dict(self._params), "states": self.states, "inputs": self.inputs}
        if self.disturbances is not None:
            out["disturbances"] = self.disturbances
        if self.name == "photobioreactor":
            out["uncertainties"] = []
        return out

    def __call__(self, x, u):
        raise NotImplementedError(
            "pcgym_b200 models are evaluated inside the CUDA kernel; there is no host RHS "
            "(use oracle.rhs in tests if you need reference values)")


class AffineCustomModel:
    """Device side of `env_params["custom_model"]` (pcgym.py:150-153).  A Python RHS cannot run inside a kernel, but
    an affine one dx = A x + B u + c can be read off the callable: c = f(0, 0), A[:, i] = f(e_i, 0) - c,
    B[:, j] = f(0, e_j) - c, then checked at random points.  The reference's own example model
    (tests/environment/test_make_env_custom_model.py:7-26) is of this form.  Anything else raises."""

    def __init__(self, user_model, rng=None):
        info = user_model.info()
        self.name = "custom_model"
        self.states, self.inputs = list(info["states"]), list(info["inputs"])
        self.disturbances = info.get("disturbances")
        if self.disturbances not in ([], None, ["None"]):
            raise NotImplementedError("custom_model with disturbance channels is not supported by the B200 engine")
        nx, nu = len(self.states), len(self.inputs)
        if not (1 <= nx <= _native.AFFINE_MAX_NX and 1 <= nu <= _native.AFFINE_MAX_NU):
            raise NotImplementedError(f"affine custom models are compiled for 1..{_native.AFFINE_MAX_NX} states and "
                                      f"1..{_native.AFFINE_MAX_NU} inputs; got {nx} x {nu}")
        import numpy as np
        f = lambda x, u: np.asarray(user_model(np.asarray(x, float), np.asarray(u, float)), dtype=np.float64).reshape(-1)
        c = f(np.zeros(nx), np.zeros(nu))
        if c.shape != (nx,):
            raise ValueError(f"custom_model returned {c.shape[0]} derivatives for {nx} states")
        A = np.stack([f(np.eye(nx)[i], np.zeros(nu)) - c for i in range(nx)], axis=1)
        B = np.stack([f(np.zeros(nx), np.eye(nu)[j]) - c for j in range(nu)], axis=1)
        rng = rng or np.random.default_rng(0)
        for _ in range(4):
            x, u = rng.normal(size=nx) * 3.0, rng.normal(size=nu) * 3.0
            want, got = f(x, u), A @ x + B @ u + c
            if not np.allclose(got, want, rtol=1e-9, atol=1e-9 * (1 + np.abs(want).max())):
                raise NotImplementedError(
                    "custom_model is not affine in (x, u): a Python right-hand side cannot run inside the CUDA "
                    "kernel; only affine custom models (dx = A x + B u + c) and the registry models are supported")
        self.A, self.B, self.c = A, B, c
        self.model_id = _native.AFFINE_BASE + (nx - 1) * _native.AFFINE_MAX_NU + (nu - 1)
        dims = _native.model_dims(self.model_id)
        assert dims[:2] == (nx, nu) and dims[3] == nx * nx + nx * nu + nx, "registry / library mismatch"

    def theta(self):
        return [float(v) for v in list(self.A.reshape(-1)) + list(self.B.reshape(-1)) + list(self.c)]


def model_mapping():
    return {name: (lambda name=name, **kw: DeviceModel(name, **kw)) for name in _META}


def get(name, int_method="b200"):
    if name in _META:
        return DeviceModel(name, int_method)
    if name in _NOT_YET:
        raise NotImplementedError(f"model '{name}' is in the reference registry but has no B200 kernel yet")
    raise ValueError(f"Model '{name}' not found in model_mapping.")   # pcgym.py:156-157
