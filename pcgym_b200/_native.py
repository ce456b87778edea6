"""ctypes binding of libpcgym_b200.so (C ABI: include/pcgym_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is
missing or does not match this binding, importing the engine raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCG_LIB") or os.path.join(HERE, "libpcgym_b200.so")     # PCG_LIB: A/B profiling knob

ABI_VERSION = 2
MAX_NX, MAX_NU, MAX_NSP, MAX_ND, MAX_NTAIL, MAX_NCON, MAX_PARAM = 24, 8, 4, 4, 8, 8, 32
AFFINE_BASE, AFFINE_MAX_NX, AFFINE_MAX_NU = 19, 4, 2
MAX_OBS = MAX_NX + MAX_NSP + MAX_ND + MAX_NTAIL

MODEL_IDS = {
    "cstr": 0,
    "first_order_system": 1,
    "multistage_extraction": 2,
    "four_tank": 3,
    "crystallization": 4,
    "biofilm_reactor": 5,
    "complex_cstr": 6,
    "nonsmooth_control": 7,
    "cstr_series_recycle": 8,
    "distillation_column": 9,
    "multistage_extraction_reactive": 10,
    "heat_exchanger": 11,
    "polymerisation_reactor": 12,
    "photobioreactor": 13,
    "invariant_batch": 14,
    "batch": 15,
    "coupled_oscillator": 16,
    "disease": 17,
    "hydraulic_tank": 18,
    "multistage_extraction_10stage": 27,       # extension (ids 19..26 are the affine custom models)
}
SCHEMES = {"rk4": 0, "dopri5": 1, "rk45": 1, "tsit5": 2, "rodas4": 3}
NOISE = {"off": 0, "philox": 1, "external": 2, "philox_f32": 3}
REWARD = {"none": 0, "sp": 1, "batch": 2, "ocp": 3}
OCP_MOMENTS = 4                     # pcg_desc.ocp_mode bit (PCG_OCP_MOMENTS)

F_NORMALISE_A, F_NORMALISE_O, F_A_DELTA, F_R_PENALTY, F_DONE_ON_VIOLATION, F_ADAPTIVE, F_SP_HIDDEN, F_EPISODE_START = (1 << i for i in range(8))
ST_NONFINITE, ST_STEPCAP, ST_RANGE = 1, 2, 4

f64, i32, u32, i64, u64 = C.c_double, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64


class Desc(C.Structure):
    """struct pcg_desc (field order and types must match the header exactly)."""
    _fields_ = [
        ("abi_version", i32), ("struct_bytes", i32), ("model", i32), ("scheme", i32), ("flags", i32),
        ("n_sub", i32), ("max_steps", i32), ("N", i32),
        ("dt", f64), ("rtol", f64), ("atol", f64),
        ("theta", f64 * MAX_PARAM),
        ("nu_a", i32), ("nd", i32), ("d_src", i32 * MAX_ND),
        ("d_default", f64 * MAX_ND),
        ("a_scale", f64 * MAX_NU), ("a_off", f64 * MAX_NU), ("a_act_lo", f64 * MAX_NU), ("a_act_hi", f64 * MAX_NU),
        ("nsp", i32), ("sp_state", i32 * MAX_NSP),
        ("r_scale", f64 * MAX_NSP), ("sp_init", f64 * MAX_NSP),
        ("nd_obs", i32), ("d_obs", i32 * MAX_ND), ("n_tail", i32), ("obs_zero_mask", u32),
        ("o_scale", f64 * MAX_OBS), ("o_off", f64 * MAX_OBS),
        ("noise", i32), ("_pad0", i32),
        ("noise_pct", f64 * MAX_NX),
        ("seed", u64), ("env_id0", i64),
        ("reward", i32), ("_pad1", i32),
        ("batch_w", f64 * MAX_NX),
        ("penalty", f64),
        ("n_con", i32), ("_pad2", i32),
        ("con_gx", (f64 * MAX_OBS) * MAX_NCON),
        ("con_gu", (f64 * MAX_NU) * MAX_NCON),
        ("con_h", f64 * MAX_NCON),
        ("con_mask_x", u64 * MAX_NCON),
        ("con_mask_u", u32 * MAX_NCON),
        ("ocp_R", f64), ("ocp_inv_span", f64 * MAX_NSP), ("ocp_mode", i32), ("_pad3", i32),
    ]


class IO(C.Structure):
    """struct pcg_io: raw device pointers (uintptr) and element strides."""
    _fields_ = [
        ("x", C.c_void_p), ("a", C.c_void_p), ("a_save", C.c_void_p),
        ("sp", C.c_void_p), ("sp_st", i64), ("sp_sj", i64), ("sp_se", i64),
        ("dist", C.c_void_p), ("d_st", i64), ("d_sj", i64), ("d_se", i64),
        ("z", C.c_void_p), ("tail", C.c_void_p),
        ("x_traj", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("violated", C.c_void_p), ("g", C.c_void_p), ("g0", C.c_void_p),
        ("done_state", C.c_void_p), ("status", C.c_void_p), ("h_carry", C.c_void_p), ("n_rhs", C.c_void_p),
        ("u_prev", C.c_void_p), ("theta", C.c_void_p),
    ]


# every symbol include/pcgym_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "pcg_abi_version": (C.c_int, []),
    "pcg_sizeof_desc": (C.c_int, []),
    "pcg_sizeof_io": (C.c_int, []),
    "pcg_model_dims": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pcg_model_stiff_capable": (C.c_int, [C.c_int]),
    "pcg_model_param_name": (C.c_char_p, [C.c_int, C.c_int]),
    "pcg_model_param_default": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "pcg_step": (C.c_int, [C.POINTER(Desc), C.POINTER(IO), C.c_int64, C.c_int64, C.c_void_p]),
    "pcg_rollout": (C.c_int, [C.POINTER(Desc), C.POINTER(IO), C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "pcg_normal_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "pcg_math_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pcg_launch_count": (C.c_int64, []),
    "pcg_strerror": (C.c_char_p, [C.c_int]),
    "pcg_last_cuda_error": (C.c_char_p, []),
    "pcg_last_kernel": (C.c_char_p, []),
    "pcg_set_kernel": (C.c_int, [C.c_int]),
}

MATH_OPS = {"exp": 0, "exp_horner": 1, "exp_tab": 2, "log": 3, "sqrt": 4, "rcp": 5, "rcp_o3": 6}
KERNELS = {"auto": -1, "generic": 0, "fast": 1, "ws": 2, "wsm": 3, "fast_ept1": 4, "generic_ept1": 5, "lanes": 6}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise NativeError if unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pcgym_b200/csrc`). pcgym_b200 has no CPU fallback.")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(L, name)
        except AttributeError as e:
            if os.environ.get("PCG_LIB") and name == "pcg_set_kernel":
                continue                    # A/B profiling against a library built before the knob existed
            raise NativeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype, fn.argtypes = res, args
    if L.pcg_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI version {L.pcg_abi_version()} != binding {ABI_VERSION}")
    if L.pcg_sizeof_desc() != C.sizeof(Desc) or L.pcg_sizeof_io() != C.sizeof(IO):
        raise NativeError(
            f"struct layout mismatch: desc {L.pcg_sizeof_desc()} vs {C.sizeof(Desc)}, io {L.pcg_sizeof_io()} vs {C.sizeof(IO)}")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        L = lib()
        msg = L.pcg_strerror(rc).decode()
        if rc == -8:
            msg += ": " + L.pcg_last_cuda_error().decode()
        raise NativeError(f"pcgym_b200 native call failed ({rc}): {msg}")


def model_dims(model_id: int):
    nx, nu, nd, npar = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().pcg_model_dims(model_id, C.byref(nx), C.byref(nu), C.byref(nd), C.byref(npar)))
    return nx.value, nu.value, nd.value, npar.value


def model_params(model_id: int) -> dict:
    """Ordered {name: default} for a registry model, read from the library."""
    L = lib()
    out = {}
    for i in range(model_dims(model_id)[3]):
        v = C.c_double()
        check(L.pcg_model_param_default(model_id, i, C.byref(v)))
        out[L.pcg_model_param_name(model_id, i).decode()] = v.value
    return out
