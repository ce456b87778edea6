#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference; the GPU box does not
have it).  Nothing here is imported by the product or by the tests; the tests
read the .npz files this script writes.

What is executed verbatim from the reference:
  * `src/pcgym/model_classes.py`  (every RHS; only `jax.numpy` is aliased to
    numpy because jax is not installed -- the models use numpy whenever
    `int_method != "jax"`).
  * `src/pcgym/pcgym.py`          (`make_env`: reset/step/rewards/constraints),
    imported under a stub `gymnasium` (Env base + Box) because gymnasium is not
    installed.
What is substituted, because CasADi / SUNDIALS cannot be installed here:
  * `pcgym.integrator.integration_engine.casadi_step` -> scipy DOP853 at
    rtol 1e-12 / atol 1e-14 on the reference's own model object (same
    zero-order-hold contract as integrator.py:90-107).

Outputs
  rhs_vectors.npz        reference RHS at random points, all 19 registry models (+ disturbance variants)
  traj_<model>.npz       golden do-mpc trajectories shipped by the reference
                         (pc-gym_paper/train_policies/*/data.npy -> ['oracle']['x','u'], rep 0)
  step_<case>.npz        per-step outputs of the reference make_env for tests/cases.py
  kat_custom_model.npz   the reference's known-answer test inputs/asserted outputs
"""
import importlib.util
import os
import sys
import types

import numpy as np
from scipy.integrate import solve_ivp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402


# ---------------------------------------------------------------- stubs ----
def _install_stubs():
    jnp = types.ModuleType("jax.numpy")
    jnp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    sys.modules["jax"], sys.modules["jax.numpy"] = jax, jnp

    class Box:
        def __init__(self, low, high, dtype=np.float64, **kw):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.dtype = dtype

        @property
        def shape(self):
            return self.low.shape

        def sample(self):
            return np.random.default_rng(0).uniform(self.low, self.high)

    gym = types.ModuleType("gymnasium")
    gym.Env = type("Env", (), {})
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = Box
    gym.spaces = spaces
    sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, spaces


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    _install_stubs()
    pkg = types.ModuleType("pcgym")
    pkg.__path__ = []
    sys.modules["pcgym"] = pkg
    mc = _load("pcgym.model_classes", f"{REF}/src/pcgym/model_classes.py")

    pe = types.ModuleType("pcgym.policy_evaluation")
    pe.policy_eval = None
    sys.modules["pcgym.policy_evaluation"] = pe

    integ = types.ModuleType("pcgym.integrator")

    class _XF:
        def __init__(self, v):
            self.v = v

        def full(self):
            return self.v

    class integration_engine:
        """scipy stand-in for integrator.py:7-107 (casadi path)."""

        def __init__(self, make_env, env_params):
            self.env = make_env(env_params)

        def casadi_step(self, state, uk):
            env = self.env
            uk = np.asarray(uk, dtype=np.float64)
            sol = solve_ivp(lambda t, x: env.model(x, uk), (0.0, env.dt), np.asarray(state[: env.Nx_oracle], float),
                            method="DOP853", rtol=1e-12, atol=1e-14)
            assert sol.success, sol.message
            return {"xf": _XF(sol.y[:, -1])}

    integ.integration_engine = integration_engine
    sys.modules["pcgym.integrator"] = integ
    env_mod = _load("pcgym.pcgym", f"{REF}/src/pcgym/pcgym.py")
    return mc, env_mod.make_env


# ---------------------------------------------------------------- RHS ----
def gen_rhs(mc, out):
    rng = np.random.default_rng(42)
    boxes = {   # (x_lo, x_hi, u_lo, u_hi[, d_lo, d_hi])
        "cstr": ([0.5, 300], [1.0, 400], [290], [305], [340, 0.9], [360, 1.1]),
        "first_order_system": ([0], [10], [0], [10]),
        "multistage_extraction": ([0.05] * 10, [0.7] * 10, [5, 10], [500, 1000], [0.5, 0.01], [0.7, 0.1]),
        "four_tank": ([0.01] * 4, [0.6] * 4, [0.1, 0.1], [10, 10]),
        "biofilm_reactor": ([0.1, 0.1, 0.1, 0.1] * 4, [10, 10, 10, 400] * 4, [0, 1, .05, .05, .05], [10, 30, 1, 1, 1]),
        "crystallization": ([1e3, 2e4, 1e6, 2e8, 0.14, 0.1, 10], [2e3, 4e4, 3e6, 4e8, 0.17, 1.5, 20], [10], [40]),
    }
    # further registry models, appended so the draws of the six above are unchanged.
    # Input-less models get one ignored input column (the device kernels carry one, csrc/models_ext.cuh).
    boxes.update({
        "complex_cstr": ([0.3, 0.0, 0.0, 300], [1.0, 0.8, 0.5, 400], [290], [305], [340, 0.9], [360, 1.1]),
        "nonsmooth_control": ([-2, -2], [2, 2], [-1], [1]),
        "cstr_series_recycle": ([20, 290, 20, 290], [100, 340, 100, 340], [1e-6, 1e-6, 280, 280], [1e-4, 1e-4, 320, 320]),
        "distillation_column": ([0.02] * 9, [0.98] * 9, [1, 150], [6, 400]),
        "multistage_extraction_reactive": ([0.05] * 20, [2.0] * 20, [5, 10], [500, 1000]),
        "heat_exchanger": ([280] * 24, [400] * 24, [0.5, 0.5, 350, 280], [5, 5, 400, 320]),
        "polymerisation_reactor": ([300, 0.5, 0.01], [400, 6.0, 0.6], [0.01, 300, 1, 0.1], [1.0, 350, 8, 1.0]),
        "photobioreactor": ([0.1, 50, 0.0], [5, 800, 0.3], [120, 0], [400, 40]),
        "invariant_batch": ([0.0] * 4, [1.0] * 4, [0], [1]),
        "batch": ([0.0, 0.0, 0.0, 280], [2.0, 1.5, 1.0, 400], [280], [350]),
        "coupled_oscillator": ([-1] * 20, [1] * 20, [0], [1]),
        "disease": ([0.0] * 3, [1.0] * 3, [0], [0.2]),
        "hydraulic_tank": ([0.0, 0.0], [2.0, 2.0], [-1], [1]),
    })
    ref_class = {"photobioreactor": "photo_production", "coupled_oscillator": "coupled_oscillators",
                 "disease": "disease_model"}
    data = {}
    M = 64
    for name, b in boxes.items():
        cls = getattr(mc, ref_class.get(name, name))
        model = cls(int_method="casadi")
        x = rng.uniform(b[0], b[1], (M, len(b[0])))
        u = rng.uniform(b[2], b[3], (M, len(b[2])))
        data[f"{name}.x"], data[f"{name}.u"] = x, u
        data[f"{name}.dx"] = np.stack([np.asarray(model(x[i], u[i]), float).reshape(-1) for i in range(M)])
        assert data[f"{name}.dx"].shape == x.shape, name
        if len(b) == 6:
            d = rng.uniform(b[4], b[5], (M, len(b[4])))
            ud = np.concatenate([u, d], axis=1)
            data[f"{name}.ud"] = ud
            data[f"{name}.dx_d"] = np.stack([cls(int_method="casadi")(x[i], ud[i]) for i in range(M)])
    np.savez_compressed(out, **data)
    print("wrote", out, len(data), "arrays")


# ---------------------------------------------------------------- golden trajectories ----
TRAJ = {   # file, dt = tsim/N (visualisation scripts)
    "cstr": ("cstr/data.npy", 26 / 60),
    "four_tank": ("four_tank/visualisation/data.npy", 1000 / 60),
    "multistage_extraction": ("multistage_extraction/data.npy", 60 / 60),
    "crystallization": ("crystalisation/data.npy", 30 / 30),
    "biofilm_reactor": ("Biofilm/data.npy", 200 / 100),
}


def gen_traj(outdir):
    for name, (rel, dt) in TRAJ.items():
        d = np.load(f"{REF}/pc-gym_paper/train_policies/{rel}", allow_pickle=True).item()["oracle"]
        x, u = d["x"][:, :, 0], d["u"][:, :, 0]
        np.savez_compressed(f"{outdir}/traj_{name}.npz", x=x, u=u, dt=np.float64(dt))
        print("wrote traj", name, x.shape, u.shape, dt)


def _script_function(path, name):
    """One function definition lifted out of a reference SCRIPT (the training scripts run SB3 on import, so they cannot
    be imported): the reference's own source text of `name`, compiled as is."""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


# ---------------------------------------------------------------- step fixtures ----
def gen_steps(make_env, outdir, only=None):
    for name, factory in cases.CASES.items():
        if only and name not in only:
            continue
        params = factory()
        if name == "cstr_ocp":      # pin the fused / restated reward against the reference's own function
            ref = _load("ref_custom_reward", f"{REF}/pc-gym_paper/train_policies/cstr/custom_reward.py")
            params["custom_reward"] = ref.sp_track_reward
        if name == "four_tank_ocp":
            params["custom_reward"] = _script_function(f"{REF}/pc-gym_paper/train_policies/four_tank/4tank_train.py", "oracle_reward")
        if name == "cryst_ocp":
            params["custom_reward"] = _script_function(f"{REF}/pc-gym_paper/train_policies/crystalisation/cryst_train.py", "oracle_reward")
        acts = cases.actions_for(name, params)
        env = make_env(params)
        draws = []
        real_normal = np.random.normal

        def recording_normal(loc=0.0, scale=1.0, size=None):
            v = real_normal(loc, scale, size)
            draws.append(np.atleast_1d(v).copy())
            return v

        np.random.seed(7)
        np.random.normal = recording_normal
        try:
            obs0, info = env.reset()
            obs0 = np.array(obs0, float)
            rec = {k: [] for k in ("obs", "reward", "done", "state", "z", "a_save", "info_obs")}
            for k in range(acts.shape[0]):
                n0 = len(draws)
                obs, r, done, trunc, info = env.step(acts[k].copy())
                z = np.full(env.Nx_oracle, np.nan)
                new = draws[n0:]
                if new:
                    if len(new) == 1 and new[0].shape[0] == env.Nx_oracle and isinstance(params.get("noise_percentage"), float):
                        z[:] = new[0]
                    else:   # per-state dict: one draw per listed state, in state order
                        names = env.model.info()["states"]
                        listed = [i for i in range(env.Nx_oracle) if names[i] in params["noise_percentage"]]
                        for i, v in zip(listed, new):
                            z[i] = v[0]
                rec["obs"].append(np.array(obs, float))
                rec["info_obs"].append(np.array(info["obs"], float))
                rec["reward"].append(float(r))
                rec["done"].append(bool(done))
                rec["state"].append(np.array(env.state, float))
                rec["z"].append(z)
                rec["a_save"].append(np.atleast_1d(np.array(getattr(env, "a_save", np.nan), float)))
                assert trunc is False
        finally:
            np.random.normal = real_normal
        out = {k: np.array(v) for k, v in rec.items()}
        out["obs0"] = obs0
        out["actions"] = acts
        if env.constraint_active:
            out["cons_info"] = np.array(info["cons_info"])
        np.savez_compressed(f"{outdir}/step_{name}.npz", **out)
        print("wrote step", name, out["obs"].shape, "final reward", out["reward"][-1], "done@", int(np.argmax(out["done"])))


def gen_kat(make_env, outdir):
    """tests/environment/test_make_env_custom_model.py:66-86 run through the reference make_env."""
    from dataclasses import dataclass, field

    @dataclass(frozen=False, kw_only=True)
    class CustomModel:
        int_method: str = field(default="casadi")
        param1: float = 1.0
        param2: float = 2.0

        def __call__(self, x, u):
            return np.array([self.param1 * x[0] + u[0], self.param2 * x[1]])

        def info(self):
            return {"parameters": {"param1": self.param1, "param2": self.param2},
                    "states": ["x1", "x2"], "inputs": ["u1"], "disturbances": []}

    env = make_env({
        "custom_model": CustomModel(param1=1.5, param2=2.5),
        "a_space": {"low": np.array([-1]), "high": np.array([1])},
        "o_space": {"low": np.array([-1, -1]), "high": np.array([1, 1])},
        "SP": {"x2": [2] * 100}, "N": 100, "tsim": 10, "x0": np.array([1.0, 1.0])})
    obs0, _ = env.reset()
    obs, r, done, _, _ = env.step(np.array([0.5]))
    np.savez(f"{outdir}/kat_custom_model.npz", obs0=obs0, obs=obs, reward=r,
             asserted=np.array([1.21578082, 1.28403262]))     # values the reference test asserts (CVODES defaults)
    print("KAT obs", obs, "asserted", [1.21578082, 1.28403262])


if __name__ == "__main__":
    mc, make_env = load_reference()
    only = sys.argv[1:]                   # e.g. `make_golden.py four_tank_ocp cryst_ocp`: just these step fixtures
    if only:
        gen_steps(make_env, HERE, only)
    else:
        gen_rhs(mc, f"{HERE}/rhs_vectors.npz")
        gen_traj(HERE)
        gen_steps(make_env, HERE)
        gen_kat(make_env, HERE)
