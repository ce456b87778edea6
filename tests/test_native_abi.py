"""The C-ABI library loads on a CPU-only box, exports every symbol include/pcgym_b200.h declares, and
rejects bad arguments with error codes (validation runs before any CUDA call)."""
import ctypes
import os
import re

import pytest

from oracle import rhs as orhs
from pcgym_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "pcgym_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pcg_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = nat.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)


def test_struct_layout_and_version():
    L = nat.lib()
    assert L.pcg_abi_version() == nat.ABI_VERSION
    assert L.pcg_sizeof_desc() == ctypes.sizeof(nat.Desc)
    assert L.pcg_sizeof_io() == ctypes.sizeof(nat.IO)
    hdr = open(os.path.join(ROOT, "include", "pcgym_b200.h")).read()
    for name, val in (("PCG_MAX_NX", nat.MAX_NX), ("PCG_MAX_NU", nat.MAX_NU), ("PCG_MAX_NSP", nat.MAX_NSP),
                      ("PCG_MAX_ND", nat.MAX_ND), ("PCG_MAX_NTAIL", nat.MAX_NTAIL), ("PCG_MAX_NCON", nat.MAX_NCON),
                      ("PCG_MAX_PARAM", nat.MAX_PARAM), ("PCG_ABI_VERSION", nat.ABI_VERSION)):
        assert int(re.search(rf"#define {name}\s+(\d+)", hdr).group(1)) == val


@pytest.mark.parametrize("model", list(nat.MODEL_IDS))
def test_model_registry_matches_oracle_metadata(model):
    mid = nat.MODEL_IDS[model]
    nx, nu, nd, npar = nat.model_dims(mid)
    meta = orhs.MODELS[model]
    assert nx == len(meta["states"]) and nu == max(len(meta["inputs"]), 1)      # input-less models: 1 ignored row
    real_d = [] if meta["disturbances"] in (["None"], []) else meta["disturbances"]
    assert nd == len(real_d)
    params = nat.model_params(mid)
    assert list(params) == list(meta["params"]) and npar == len(params)
    for k, v in meta["params"].items():
        assert params[k] == v


def _valid_desc():
    d = nat.Desc()
    d.abi_version, d.struct_bytes = nat.ABI_VERSION, ctypes.sizeof(nat.Desc)
    d.model, d.scheme, d.n_sub, d.max_steps, d.N, d.dt = 0, 2, 1, 100, 10, 0.1
    d.rtol, d.atol, d.nu_a = 1e-8, 1e-8, 1
    return d


def test_argument_validation_codes():
    L = nat.lib()
    io = nat.IO()
    call = lambda d, n=8, t=0, k=1: L.pcg_rollout(ctypes.byref(d), ctypes.byref(io), n, t, k, None)
    assert L.pcg_rollout(None, ctypes.byref(io), 8, 0, 1, None) == -1
    d = nat.Desc()
    assert call(d) == -2                                   # version / size
    d = _valid_desc(); d.model = 99
    assert call(d) == -3
    d = _valid_desc(); d.scheme = 7
    assert call(d) == -4
    d = _valid_desc(); d.scheme = 0; d.flags = nat.F_ADAPTIVE
    assert call(d) == -4                                   # adaptive RK4
    d = _valid_desc(); d.scheme = nat.SCHEMES["rodas4"]
    assert call(d) == -4                                   # the Rosenbrock scheme is step-size controlled only
    d = _valid_desc(); d.scheme = nat.SCHEMES["rodas4"]; d.flags = nat.F_ADAPTIVE; d.model = nat.MODEL_IDS["four_tank"]; d.nu_a = 2
    assert call(d) == -4                                   # ... and needs a model with an analytic Jacobian
    assert L.pcg_model_stiff_capable(nat.MODEL_IDS["cstr"]) == 1 and L.pcg_model_stiff_capable(nat.MODEL_IDS["multistage_extraction"]) == 1
    assert L.pcg_model_stiff_capable(nat.MODEL_IDS["four_tank"]) == 0 and L.pcg_model_stiff_capable(99) == 0
    d = _valid_desc(); d.nu_a = 3
    assert call(d) == -5
    d = _valid_desc(); d.nd = 1
    assert call(d) == -5                                   # cstr has 2 disturbance channels: 0 or 2
    d = _valid_desc(); d.nsp = 1; d.sp_state[0] = 5
    assert call(d) == -5
    d = _valid_desc(); d.n_sub = 0
    assert call(d) == -6
    d = _valid_desc()
    assert call(d, n=0) == -6 and call(d, k=0) == -6 and call(d, t=-1) == -6
    assert call(d, t=8, k=2) == -6 and call(d, t=0, k=10) == -6      # t0 + k_steps must stay within N - 1 = 9 intervals
    assert call(d) == -1                                   # x / a pointers missing
    io.x, io.a = 0x1004, 0x2000
    assert call(d) == -7                                   # misaligned
    for rc in range(0, -9, -1):
        assert L.pcg_strerror(rc) not in (None, b"unknown status")
    assert L.pcg_step(ctypes.byref(_valid_desc()), ctypes.byref(nat.IO()), 8, 0, None) == -1
    # the diagnostic entry points validate before they launch as well
    assert L.pcg_math_probe(nat.MATH_OPS["log"], None, None, 8, None) == -1
    assert L.pcg_math_probe(99, 0x1000, 0x2000, 8, None) == -6 and L.pcg_math_probe(0, 0x1000, 0x2000, 0, None) == -6
    assert L.pcg_normal_pairs(None, None, None, None, 8, 0, None) == -1
    assert L.pcg_launch_count() == 0                       # nothing was launched by any of the above
    with pytest.raises(nat.NativeError):
        nat.check(-3)


def test_kernel_request_knob():
    L = nat.lib()
    assert L.pcg_set_kernel(nat.KERNELS["wsm"]) == nat.KERNELS["auto"]          # returns the previous request
    assert L.pcg_set_kernel(nat.KERNELS["auto"]) == nat.KERNELS["wsm"]
    assert L.pcg_set_kernel(7) == -6 and L.pcg_set_kernel(-2) == -6
    assert L.pcg_set_kernel(nat.KERNELS["lanes"]) == nat.KERNELS["auto"] and L.pcg_set_kernel(nat.KERNELS["auto"]) == nat.KERNELS["lanes"]
    assert L.pcg_set_kernel(nat.KERNELS["auto"]) == nat.KERNELS["auto"]
