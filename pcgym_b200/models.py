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

# registry keys of the reference that have no device RHS yet
_NOT_YET = ["complex_cstr", "nonsmooth_control", "cstr_series_recycle", "distillation_column",
            "multistage_extraction_reactive", "heat_exchanger", "polymerisation_reactor", "photobioreactor",
            "invariant_batch", "batch", "coupled_oscillator", "disease", "hydraulic_tank"]


class DeviceModel:
    """A registry model whose RHS runs in the CUDA library."""

    def __init__(self, name: str, int_method: str = "b200", **overrides):
        if name not in _META:
            raise KeyError(name)
        self.name = name
        self.model_id = _native.MODEL_IDS[name]
        self.int_method = int_method
        self.states, self.inputs, self.disturbances = (list(v) for v in _META[name])
        params = _native.model_params(self.model_id)
        for k, v in overrides.items():
            if k not in params:
                raise AttributeError(f"{name} has no parameter {k!r}")
            params[k] = float(v)
        self.__dict__["_params"] = params
        nx, nu, nd, _ = _native.model_dims(self.model_id)
        real_nd = 0 if self.disturbances in (["None"], []) else len(self.disturbances)
        assert (nx, nu, nd) == (len(self.states), len(self.inputs), real_nd), "registry / library mismatch"

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
        return {"parameters": dict(self._params), "states": self.states, "inputs": self.inputs,
                "disturbances": self.disturbances}

    def __call__(self, x, u):
        raise NotImplementedError(
            "pcgym_b200 models are evaluated inside the CUDA kernel; there is no host RHS "
            "(use oracle.rhs in tests if you need reference values)")


def model_mapping():
    return {name: (lambda name=name, **kw: DeviceModel(name, **kw)) for name in _META}


def get(name, int_method="b200"):
    if name in _META:
        return DeviceModel(name, int_method)
    if name in _NOT_YET:
        raise NotImplementedError(f"model '{name}' is in the reference registry but has no B200 kernel yet")
    raise ValueError(f"Model '{name}' not found in model_mapping.")   # pcgym.py:156-157
