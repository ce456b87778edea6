"""`make_env`: the reference's Gymnasium surface (src/pcgym/pcgym.py:31-500) over the B200 kernels.

Drop-in contract
  * same constructor argument (`env_params` dict, same 31 keys), same attributes
    (N, tsim, dt, Nx, Nx_oracle, Nu, Nd, Nd_model, SP, x0, t, state, obs, done, info, model, a_save, ...),
    same `reset(seed=0) -> (obs, info)` and `step(a) -> (obs, r, done, False, info)`;
  * `n_envs` (default 1) independent copies advance per launch.  With `n_envs == 1` and numpy input
    the return types are the reference's (float64 ndarray / float / bool); otherwise torch CUDA
    tensors: obs [n_envs, Nx], reward [n_envs], done [n_envs] bool, truncated [n_envs] bool;
  * `rollout(actions)` advances K control intervals in ONE launch (state stays in registers).

Everything numeric on the step path happens inside `libpcgym_b200.so` (pcg_step / pcg_rollout); this
file parses the configuration into a `pcg_desc`, owns the device tensors (struct-of-arrays
[field][env], fp64) and mirrors the reference's bookkeeping.  There is no CPU fallback.

New optional keys: `n_envs`, `device`, `seed`, `env_id0`, `noise_mode` ("philox" | "philox_f32" |
"external"), `integrator` = {"scheme": "tsit5"|"dopri5"|"rk4"|"rodas4" (stiff: Rosenbrock 4(3) with the model's
analytic Jacobian, cstr and multistage_extraction), "adaptive": bool, "rtol", "atol", "n_sub", "max_steps"}, `constraint_units` ("reference" | "physical"), per-env `x0` [n_envs, Nx],
per-env `SP[k]` / `disturbances[k]` [n_envs, N]; per-env model parameters that reach the DYNAMICS (an
extension: the reference's sampled parameters only reach the observation): `model_params` =
{"k0": array[n_envs], ...} and/or `uncertainty_in_dynamics=True`, which feeds the values reset() samples for
`uncertainty_percentages` / `empirical_distribution` to the kernel as well.
"""
from __future__ import annotations

import copy
import ctypes

import warnings

import numpy as np
import torch

from . import _native as nat
from . import models as _models
from .spaces import Box

_DEFAULT_INTEGRATOR = {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8, "n_sub": 1,
                       "max_steps": 4096}     # mirrors the reference's jax path: Tsit5 + PID(1e-8, 1e-8), integrator.py:57-61


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy().astype(np.float64)
    return np.asarray(a, dtype=np.float64)


class make_env:
    metadata = {"render_modes": []}

    # ------------------------------------------------------------------ construction (host only)
    def __init__(self, env_params: dict) -> None:
        if not isinstance(env_params, dict):
            raise ValueError("env_params must be a dictionary")                    # pcgym.py:40-41
        nat.lib()                                                                  # fail loudly if the .so is absent
        P = self.env_params = copy.deepcopy(env_params)
        self.n_envs = int(P.get("n_envs", 1))
        if self.n_envs < 1:
            raise ValueError("n_envs must be >= 1")
        self.device = torch.device(P.get("device", "cuda"))
        self.seed = int(P.get("seed", 0))
        self.env_id0 = int(P.get("env_id0", 0))
        self._copy_outputs = bool(P.get("copy_outputs", True))

        # pcgym.py:56-61
        self.a_delta = P.get("a_delta", False)
        if self.a_delta:
            self.a_0 = P["a_0"]
        self.normalise_a = P.get("normalise_a", True)
        self.normalise_o = P.get("normalise_o", True)

        # pcgym.py:68-92
        self._a_lo, self._a_hi = _np(P["a_space"]["low"]).ravel(), _np(P["a_space"]["high"]).ravel()
        n_a = self._a_lo.shape[0]
        self.action_space = Box(-np.ones(n_a), np.ones(n_a)) if self.normalise_a else Box(self._a_lo, self._a_hi)
        o_lo, o_hi = _np(P["o_space"]["low"]).ravel(), _np(P["o_space"]["high"]).ravel()

        # pcgym.py:94-111
        self.maximise_reward = True
        self.SP = P.get("SP")
        self.custom_reward = P.get("custom_reward") is not None
        self._ocp = None
        if self.custom_reward:
            self.custom_reward_f = P["custom_reward"]
            self.reward = None
            if isinstance(self.custom_reward_f, dict):
                # declarative form of the paper scripts' reward: fused into the kernel (PCG_REWARD_OCP)
                # kind "ocp_tracking": custom_reward.py:3-39 / 4tank_train.py:16-52; kind "cryst_tracking": the
                # crystallization script's variant (cryst_train.py:17-48), which tracks CV and Ln DERIVED from the
                # noisy moments against the first and second setpoint (R defaults to that script's 0.01)
                spec = dict(self.custom_reward_f)
                kind = spec.pop("kind", "ocp_tracking")
                if kind not in ("ocp_tracking", "cryst_tracking"):
                    raise ValueError("custom_reward dict: kind must be 'ocp_tracking' or 'cryst_tracking'")
                self._ocp = {"R": float(spec.pop("R", 0.1 if kind == "ocp_tracking" else 0.01)),
                             "delta_u": bool(spec.pop("delta_u", True)), "moments": kind == "cryst_tracking"}
                if self._ocp["moments"] and (P.get("model") != "crystallization" or len(self.SP or {}) != 2):
                    raise ValueError("custom_reward kind='cryst_tracking' is the crystallization model's reward "
                                     "(two setpoints: CV, Ln)")
                if spec:
                    raise ValueError(f"custom_reward dict: unknown keys {sorted(spec)}")
                if self.SP is None:
                    raise ValueError("custom_reward kind='ocp_tracking' needs SP")
            elif not callable(self.custom_reward_f):
                raise ValueError("custom_reward must be a callable or an 'ocp_tracking' dict")
        elif self.SP is not None:
            self.reward = "SP_reward_fn"
        else:
            self.reward = "batch_reward_fn"
            self.reward_states = P["reward_states"]
            self.maximise_reward = P["maximise_reward"]
        self.N = int(P["N"])
        self.tsim = P["tsim"]
        self.x0 = P["x0"]
        self.integration_method = P.get("integration_method", "b200")
        if self.integration_method not in ("b200", "casadi", "jax"):              # integrator.py:33-36
            raise AssertionError("integration_method must be 'b200' (or the reference's 'casadi' / 'jax', which map to it)")
        self.dt = self.tsim / self.N
        self.done = False
        integ = dict(_DEFAULT_INTEGRATOR)
        integ.update(P.get("integrator", {}))
        if integ["scheme"] not in nat.SCHEMES:
            raise ValueError(f"unknown integrator scheme {integ['scheme']!r}")
        if integ["scheme"] == "rk4":
            integ["adaptive"] = False
        if integ["scheme"] == "rodas4":               # Rosenbrock 4(3): always step-size controlled
            integ["adaptive"] = True
        self.integrator = integ
        self._fixed_checked = False
        self._episode_fresh = False
        self.fixed_step_probe_error = None       # set by the first reset() of a fixed-step env (fixed_step_error)

        # pcgym.py:127-158
        if P.get("custom_model") is not None:
            # env.model stays the user's object (pcgym.py:150-153); its affine RHS is read off it for the device
            m = P["custom_model"]
            self._dmodel = _models.AffineCustomModel(m)          # raises for a non-affine right-hand side
            m.int_method = self.integration_method
            self.model = m
            if self.integrator["scheme"] != "tsit5":
                raise ValueError("custom (affine) models run with the 'tsit5' integrator scheme")
        else:
            self.model = self._dmodel = _models.get(P.get("model"), self.integration_method)
        if self.integrator["scheme"] == "rodas4" and not nat.lib().pcg_model_stiff_capable(self._dmodel.model_id):
            raise ValueError(f"integrator scheme 'rodas4' needs the model's analytic Jacobian: available for 'cstr' and "
                             f"'multistage_extraction', not for '{self._dmodel.name}'")
        minfo = self.model.info()

        # pcgym.py:160-165
        self.Nx_oracle = len(minfo["states"])
        self.Nx = self.Nx_oracle + (len(self.SP) if self.SP is not None else 0)
        # input-less models (invariant_batch, coupled_oscillator) take one ignored action entry (csrc/models_ext.cuh)
        self.Nu = max(len(minfo["inputs"]), 1)
        if n_a != self.Nu:
            raise ValueError(f"a_space has {n_a} entries but model '{self._dmodel.name}' has {self.Nu} inputs")

        # pcgym.py:167-199
        self.disturbance_active = False
        self.Nd = self.Nd_model = 0
        if P.get("disturbances") is not None:
            self.disturbance_active = True
            self.disturbances = P["disturbances"]
            self.Nd = len(self.disturbances)
            self.Nd_model = len(minfo.get("disturbances") or [])
            for k in (minfo.get("disturbances") or []):
                if k not in self.disturbances and str(k) not in minfo["parameters"]:
                    raise KeyError(str(k))                                         # pcgym.py:400 on the reference
            unknown = [k for k in self.disturbances if k not in (minfo.get("disturbances") or [])]
            if unknown:
                raise KeyError(f"disturbances {unknown} are not channels of model '{self._dmodel.name}' "
                               f"({minfo['disturbances']})")
            self.Nu += self.Nd_model
            self.Nx += self.Nd
            o_lo = np.concatenate((o_lo, _np(P["disturbance_bounds"]["low"]).ravel()))
            o_hi = np.concatenate((o_hi, _np(P["disturbance_bounds"]["high"]).ravel()))

        # pcgym.py:212-252 (observation-only in the reference: sampled values never reach the dynamics)
        self.uncertainty = False
        self.NUn = 0
        self.uncertainty_percentages = None
        self.empirical_distribution = None
        if P.get("uncertainty_percentages") is not None or P.get("empirical_distribution") is not None:
            self.uncertainty = True
            if P.get("uncertainty_percentages") is not None:
                self.uncertainty_percentages = P["uncertainty_percentages"]
                names = [k for k in self.uncertainty_percentages if k != "x0"]
                self.distribution = P.get("distribution")
            else:
                self.empirical_distribution = P["empirical_distribution"]
                names = [k for k in self.empirical_distribution if k != "x0"]
            self.original_param_values = {k: getattr(self.model, k) for k in names}
            self._unc_names = names
            self.NUn = len(names)
            o_lo = np.concatenate((o_lo, _np(P["uncertainty_bounds"]["low"]).ravel()))
            o_hi = np.concatenate((o_hi, _np(P["uncertainty_bounds"]["high"]).ravel()))
        # extension: parameters that differ per env and DO reach the dynamics
        self.uncertainty_in_dynamics = bool(P.get("uncertainty_in_dynamics", False)) and self.uncertainty
        self._theta_names = list(self.model.info()["parameters"])
        self._theta_user = {}
        for k, v in (P.get("model_params") or {}).items():
            if k not in self._theta_names:
                raise AttributeError(f"{self._dmodel.name} has no parameter {k!r}")
            v = _np(v).reshape(-1)
            if v.shape[0] not in (1, self.n_envs):
                raise ValueError(f"model_params[{k!r}] must have n_envs = {self.n_envs} entries")
            self._theta_user[k] = np.broadcast_to(v, (self.n_envs,)).copy()
        self._per_env_theta = bool(self._theta_user) or self.uncertainty_in_dynamics
        if self._per_env_theta and self._dmodel is not self.model:
            raise NotImplementedError("per-env parameters are not supported for custom models")
        self.observation_space_base = Box(o_lo, o_hi)
        n_o = o_lo.shape[0]
        self.observation_space = Box(-np.ones(n_o), np.ones(n_o)) if self.normalise_o else self.observation_space_base
        self._n_obs = self.Nx + self.NUn
        # The reference never checks len(x0): with x0 / o_space that carry NO setpoint slots the setpoint write
        # (pcgym.py:437) falls off the end of the state vector and the observation is [x] only, while the reward
        # still tracks SP -- its own custom-model test runs like this (test_make_env_custom_model.py:66-86).
        n_sp = len(self.SP) if self.SP is not None else 0
        x0_len = _np(self.x0).shape[-1]
        self._sp_hidden = bool(n_sp and x0_len == self.Nx_oracle and n_o == self._n_obs - n_sp)
        if self._sp_hidden:
            if self.Nd or self.NUn:
                raise NotImplementedError("x0 without setpoint slots together with observed disturbances / uncertain "
                                          "parameters is not supported")
            self._n_obs -= n_sp
        if n_o != self._n_obs:
            raise ValueError(f"o_space (+ disturbance/uncertainty bounds) has {n_o} entries, state vector has {self._n_obs}")

        # pcgym.py:63-66, 208-211
        self.noise_percentage = P.get("noise_percentage")
        if self.noise_percentage is not None:
            self.noise_percentage_float = isinstance(self.noise_percentage, float)
        self.partial_observation = P.get("partial_observation") or False
        self.noise_mode = P.get("noise_mode", "philox")
        if self.noise_mode not in ("philox", "philox_f32", "external"):
            raise ValueError("noise_mode must be 'philox', 'philox_f32' or 'external'")

        # pcgym.py:113-125 (after the model so the affine probe knows the dimensions)
        self.constraint_active = False
        self._host_constraints = False
        self.r_penalty = False
        self.info = {}
        self.n_con = 0
        if P.get("constraints") is not None:
            self.constraints = P["constraints"]
            self.done_on_constraint = P["done_on_cons_vio"]
            self.r_penalty = P["r_penalty"]
            self.constraint_active = True
            # A callable that is not affine in (state, uk) cannot be fused: it then runs UNFUSED in step() on the
            # tensors the kernel returns (like a Python custom_reward), and rollout() refuses it.
            self._host_constraints = False
            try:
                self._G, self._h = self._affine_constraints()
                self.n_con = self._h.shape[0]
                if self.n_con > nat.MAX_NCON:
                    raise ValueError(f"at most {nat.MAX_NCON} constraints are supported")
            except NotImplementedError:
                self._host_constraints = True
                probe = np.asarray(self.constraints(np.ones(self._n_obs), np.ones(self.Nu)), dtype=np.float64)
                self.n_con = int(probe.reshape(-1).shape[0])                       # pcgym.py:124

        self._host_tables()
        self._desc = self._build_desc()
        # step() returns info["obs"] BEFORE the partial-observation mask (pcgym.py:489-498), so its launches run with
        # the mask off and the mask is applied afterwards; rollout() applies it in the kernel
        self._desc_step = self._desc
        if self._desc.obs_zero_mask:
            self._desc_step = nat.Desc.from_buffer_copy(self._desc)
            self._desc_step.obs_zero_mask = 0
        self._dev_ready = False
        self.t = 0
        # Episode counter: the reference draws fresh np.random values every step and every reset (pcgym.py:257-260,
        # 458, 464).  Here the noise is a pure function of (key, env, t), so the key advances with the episode:
        # key = seed + episode * 0x9E3779B97F4A7C15 (mod 2^64); reset() / rewind() start the next episode,
        # reset(episode=k) / rewind(episode=k) pin it (reproducible runs, CUDA-graph capture).
        self.episode = -1

    # ------------------------------------------------------------------ constraints: callable -> affine form
    def _affine_constraints(self):
        """The reference takes a Python callable g(x, u) <= 0 (pcgym.py:571-572).  A callable cannot run
        in the kernel, so it is probed: g(0), g(e_i) give h and the columns of G, and random points verify
        that g really is affine in (state, uk).  Every constraint in the reference's docs and tests is."""
        nx, nu = self._n_obs, self.Nu
        if isinstance(self.constraints, dict):
            return self._box_constraints()
        g = lambda s, u: np.asarray(self.constraints(s, u), dtype=np.float64).reshape(-1)
        h = g(np.zeros(nx), np.zeros(nu))
        G = np.zeros((h.shape[0], nx + nu))
        for i in range(nx + nu):
            e = np.zeros(nx + nu)
            e[i] = 1.0
            G[:, i] = g(e[:nx], e[nx:]) - h
        rng = np.random.default_rng(0)
        s0 = np.ones(nx)                                   # representative magnitudes for the affinity check
        x0 = np.abs(_np(self.x0).reshape(-1, _np(self.x0).shape[-1])[0])
        s0[: min(nx, x0.shape[0])] = x0[:nx]
        u0 = np.ones(nu)
        u0[: self._a_hi.shape[0]] = np.maximum(np.abs(self._a_hi), np.abs(self._a_lo))
        scale = np.concatenate((np.where(s0 > 0, s0, 1.0), np.where(u0 > 0, u0, 1.0)))
        for _ in range(4):
            v = rng.normal(size=nx + nu) * scale
            want, got = g(v[:nx], v[nx:]), G @ v + h
            if not np.allclose(want, got, rtol=1e-9, atol=1e-9 * (1 + np.abs(want).max())):
                raise NotImplementedError(
                    "constraints(x, u) is not affine in (x, u): only affine constraints can be fused "
                    "into the B200 step kernel")
        return G, h

    def _box_constraints(self):
        """Declarative bounds, the API of the reference's oracle / paper scripts
        (`constraints={'T': [327, 321]}, cons_type={'T': ['<=', '>=']}`, oracle.py:127-134,
        pc-gym_paper/constraint_showcase/constraint_paper.py:33-35).  Physical units; one row per bound."""
        names = self.model.info()["states"]
        inputs = self.model.info()["inputs"]
        ctype = self.env_params.get("cons_type")
        if ctype is None:
            raise KeyError("cons_type")
        rows, hs = [], []
        for k, bounds in self.constraints.items():
            for b, op in zip(np.atleast_1d(bounds), ctype[k]):
                row = np.zeros(self._n_obs + self.Nu)
                if k in names:
                    col = names.index(k)
                elif k in inputs:
                    col = self._n_obs + inputs.index(k)
                else:
                    raise ValueError(f"'{k}' is not in list")
                if op == "<=":
                    row[col], h = 1.0, -float(b)            # x - b <= 0
                elif op == ">=":
                    row[col], h = -1.0, float(b)            # b - x <= 0
                else:
                    raise ValueError(f"cons_type entries must be '<=' or '>=', got {op!r}")
                rows.append(row)
                hs.append(h)
        self.env_params.setdefault("constraint_units", "physical")
        return np.array(rows), np.array(hs)

    # ------------------------------------------------------------------ host tables (setpoints, disturbances, x0)
    def _per_time(self, v, what):
        """-> (array [N] shared) or (array [n_envs, N] per env)."""
        a = _np(v)
        if a.ndim == 1:
            if a.shape[0] < self.N:
                raise IndexError(f"{what} has {a.shape[0]} entries, needs N = {self.N}")
            return a[: self.N], False
        if a.shape != (self.n_envs, a.shape[1]) or a.shape[1] < self.N:
            raise ValueError(f"{what} must be [N] or [n_envs, N]")
        return a[:, : self.N], True

    def _host_tables(self):
        names = self.model.info()["states"]
        # setpoints, dict order (pcgym.py:432-438)
        self._sp_keys = list(self.SP.keys()) if self.SP is not None else []
        if len(self._sp_keys) > nat.MAX_NSP:
            raise ValueError(f"at most {nat.MAX_NSP} setpoints are supported")
        for k in self._sp_keys:
            if k not in names:
                raise ValueError(f"'{k}' is not in list")                          # names.index(k), pcgym.py:553
        sp = [self._per_time(self.SP[k], f"SP['{k}']") for k in self._sp_keys]
        self._sp_per_env = any(pe for _, pe in sp)
        if self._sp_keys:
            if self._sp_per_env:     # [N, nsp, n_envs]
                self._sp_host = np.stack([np.broadcast_to(a if pe else a[None], (self.n_envs, self.N)).T
                                          for a, pe in sp], axis=1)
            else:                    # [nsp, N]
                self._sp_host = np.stack([a for a, _ in sp])
        else:
            self._sp_host = None
        # disturbances in model order (pcgym.py:392-405)
        self._d_keys = [k for k in (self.model.info().get("disturbances") or []) if self.disturbance_active and k in self.disturbances]
        ds = [self._per_time(self.disturbances[k], f"disturbances['{k}']") for k in self._d_keys]
        self._d_per_env = any(pe for _, pe in ds)
        if self._d_keys:
            if self._d_per_env:
                self._d_host = np.stack([np.broadcast_to(a if pe else a[None], (self.n_envs, self.N)).T
                                         for a, pe in ds], axis=1)
            else:
                self._d_host = np.stack([a for a, _ in ds])
        else:
            self._d_host = None
        # x0: [Nx_full] or [n_envs, Nx_full]; only the first Nx_oracle + nSP entries are used (pcgym.py:284)
        x0 = _np(self.x0)
        if x0.ndim == 1:
            x0 = np.broadcast_to(x0, (self.n_envs, x0.shape[0]))
        elif x0.shape[0] != self.n_envs:
            raise ValueError("per-env x0 must be [n_envs, Nx]")
        need = self.Nx_oracle + (0 if self._sp_hidden else len(self._sp_keys))
        if x0.shape[1] != need:
            raise ValueError(f"x0 has {x0.shape[1]} entries; expected {need} (model states + setpoints)")
        self._x0_host = np.ascontiguousarray(x0.T)        # [Nx_oracle + nsp, n_envs]

    # ------------------------------------------------------------------ descriptor
    def _build_desc(self) -> nat.Desc:
        d = nat.Desc()
        P = self.env_params
        nx = self.Nx_oracle
        names = self.model.info()["states"]
        d.abi_version, d.struct_bytes = nat.ABI_VERSION, ctypes.sizeof(nat.Desc)
        d.model = self._dmodel.model_id
        d.scheme = nat.SCHEMES[self.integrator["scheme"]]
        flags = 0
        flags |= nat.F_NORMALISE_A if self.normalise_a is True else 0
        self._host_reward = self.custom_reward and self._ocp is None      # Python callable: unfused path
        flags |= nat.F_NORMALISE_O if (self.normalise_o is True and not self._host_reward) else 0
        flags |= nat.F_A_DELTA if (self.a_delta and self.normalise_a) else 0
        fused_con = self.constraint_active and not self._host_constraints
        flags |= nat.F_R_PENALTY if (self.r_penalty and fused_con) else 0
        flags |= nat.F_DONE_ON_VIOLATION if (fused_con and self.done_on_constraint) else 0
        flags |= nat.F_ADAPTIVE if self.integrator["adaptive"] else 0
        flags |= nat.F_SP_HIDDEN if self._sp_hidden else 0
        d.flags = flags
        d.n_sub, d.max_steps = int(self.integrator["n_sub"]), int(self.integrator["max_steps"])
        d.rtol, d.atol = float(self.integrator["rtol"]), float(self.integrator["atol"])
        d.N, d.dt = self.N, float(self.dt)
        theta = self._dmodel.theta()
        for i, v in enumerate(theta):
            d.theta[i] = v

        nu_a = self._a_lo.shape[0]
        d.nu_a = nu_a
        params = self.model.info()["parameters"]
        model_d = [] if self._dmodel.disturbances in (["None"], [], None) else self._dmodel.disturbances
        d.nd = len(model_d) if self.disturbance_active else 0
        for i in range(nat.MAX_ND):
            d.d_src[i] = -1
        for i, k in enumerate(model_d):
            d.d_default[i] = float(params.get(k, 0.0))
            if self.disturbance_active and k in self._d_keys:
                d.d_src[i] = self._d_keys.index(k)
        for j in range(nu_a):
            half = (self._a_hi[j] - self._a_lo[j]) / 2.0
            d.a_scale[j], d.a_off[j] = half, self._a_lo[j] + half
            if self.a_delta:
                d.a_act_lo[j] = float(_np(P["a_space_act"]["low"]).ravel()[j])
                d.a_act_hi[j] = float(_np(P["a_space_act"]["high"]).ravel()[j])

        d.nsp = len(self._sp_keys)
        r_scale = P.get("r_scale", {})
        for j, k in enumerate(self._sp_keys):
            d.sp_state[j] = names.index(k)
            d.r_scale[j] = float(r_scale.get(k, 1))
            d.sp_init[j] = 0.0 if self._sp_hidden else float(self._x0_host[nx + j, 0])
        d.nd_obs = len(self._d_keys)
        for j, k in enumerate(self._d_keys):
            d.d_obs[j] = model_d.index(k)
        d.n_tail = self.NUn
        lo, hi = self.observation_space_base.low.astype(np.float64), self.observation_space_base.high.astype(np.float64)
        for i in range(self._n_obs):
            span = hi[i] - lo[i]
            d.o_scale[i] = 2.0 / span if span != 0 else float("inf")
            d.o_off[i] = -2.0 * lo[i] / span - 1.0 if span != 0 else float("nan")
        mask = 0
        if self.partial_observation:
            for i in range(nx):
                if names[i] not in self.partial_observation:
                    mask |= 1 << i
        self._obs_mask_bits = mask
        self._mask_rows = [i for i in range(nx) if mask >> i & 1]
        d.obs_zero_mask = 0 if self._host_reward else mask

        d.noise = nat.NOISE["off"]
        if P.get("noise", False):
            if self.noise_percentage is None:
                raise AttributeError("noise=True needs noise_percentage")          # pcgym.py:455 on the reference
            d.noise = nat.NOISE[self.noise_mode]
            for i in range(nx):
                if self.noise_percentage_float:
                    d.noise_pct[i] = float(self.noise_percentage)
                elif names[i] in self.noise_percentage:
                    d.noise_pct[i] = float(self.noise_percentage[str(names[i])])
        d.seed, d.env_id0 = self.seed & (2 ** 64 - 1), self.env_id0

        d.penalty = 1000.0
        if self._ocp is not None:
            d.reward = nat.REWARD["ocp"]
            d.ocp_R = self._ocp["R"]
            d.ocp_mode = (2 if self._ocp["delta_u"] else 0) | (nat.OCP_MOMENTS if self._ocp["moments"] else 0)
            for j, k in enumerate(self._sp_keys):
                i = names.index(k)
                d.ocp_inv_span[j] = 1.0 / (hi[i] - lo[i])
        elif self.custom_reward:
            d.reward = nat.REWARD["none"]
        elif self.reward == "SP_reward_fn":
            d.reward = nat.REWARD["sp"]
        else:
            d.reward = nat.REWARD["batch"]
            for s in self.reward_states:                                            # pcgym.py:519-527
                if str(s) in names:
                    w = float(r_scale.get(s, 1))
                    d.batch_w[names.index(s)] += w if self.maximise_reward else -w

        d.n_con = self.n_con if fused_con else 0
        if fused_con:
            G, h = self._G.copy(), self._h.copy()
            n_s = self._n_obs
            if P.get("constraint_units", "reference") == "reference":
                # constraint_check un-normalises values that are already physical (pcgym.py:597-608);
                # both maps are affine, so the quirk folds into G and h.
                if self.normalise_a is True:
                    if nu_a not in (1, self.Nu):
                        raise ValueError(f"operands could not be broadcast together with shapes ({self.Nu},) ({nu_a},)")
                    sa = np.broadcast_to((self._a_hi - self._a_lo) / 2.0, (self.Nu,))
                    la = np.broadcast_to(self._a_lo, (self.Nu,))
                    h = h + G[:, n_s:] @ (sa + la)
                    G[:, n_s:] = G[:, n_s:] * sa
                if self.normalise_o is True:
                    so = (hi - lo) / 2.0
                    h = h + G[:, :n_s] @ (so + lo)
                    G[:, :n_s] = G[:, :n_s] * so
            for c in range(self.n_con):
                mx = mu = 0
                for i in range(n_s):
                    d.con_gx[c][i] = G[c, i]
                    if G[c, i] != 0.0:
                        mx |= 1 << i
                for i in range(self.Nu):
                    # uk on the device always carries the model's disturbance channels after the actions
                    d.con_gu[c][i] = G[c, n_s + i]
                    if G[c, n_s + i] != 0.0:
                        mu |= 1 << i
                d.con_h[c] = h[c]
                d.con_mask_x[c], d.con_mask_u[c] = mx, mu
        return d

    # ------------------------------------------------------------------ device state
    def _ensure_device(self):
        if self._dev_ready:
            return
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("pcgym_b200 needs a CUDA device (no CPU fallback); "
                               f"device={self.device}, cuda available={torch.cuda.is_available()}")
        dev, n, f64 = self.device, self.n_envs, torch.float64
        nx = self.Nx_oracle
        z = lambda *s, dtype=f64: torch.zeros(*s, dtype=dtype, device=dev)
        self._x = z(nx, n)
        self._x0_dev = torch.from_numpy(np.array(self._x0_host)).to(dev)
        self._sp_dev = torch.from_numpy(np.ascontiguousarray(self._sp_host)).to(dev) if self._sp_host is not None else None
        self._d_dev = torch.from_numpy(np.ascontiguousarray(self._d_host)).to(dev) if self._d_host is not None else None
        self._a_save = z(self._a_lo.shape[0], n) if self.a_delta else None
        self._u_prev = z(self._a_lo.shape[0], n) if self._ocp is not None else None
        self._tail = z(max(self.NUn, 1), n) if self.NUn else None
        self._done_state = z(n, dtype=torch.uint8)
        self._status = z(n, dtype=torch.uint8)
        self._h_carry = z(n)
        self._n_rhs = z(n, dtype=torch.int32)
        # cons_info lives as [N, n_con, n_envs] so the kernel writes rows in place; exposed as (n_con, N, n_envs)
        self._cons = z(self.N, max(self.n_con, 1), n) if self.n_con else None      # (host constraints write it from step())
        self._theta = None
        if self._per_env_theta:                      # [nparam, n]: defaults, then the user's per-env rows
            base = torch.tensor([float(v) for v in self.model.info()["parameters"].values()], dtype=f64, device=dev)
            self._theta = base[:, None].expand(-1, n).contiguous()
            for k, v in self._theta_user.items():
                self._theta[self._theta_names.index(k)] = torch.from_numpy(v).to(dev)
        self._bufs = {}
        self._trunc = None
        self._pipe = None
        self._pipe_prev = None
        self._a_step = z(1, self._a_lo.shape[0], n)
        self._o_lo = torch.from_numpy(self.observation_space_base.low.astype(np.float64)).to(dev)
        self._o_hi = torch.from_numpy(self.observation_space_base.high.astype(np.float64)).to(dev)
        self._dev_ready = True

    def _out_bufs(self, K, want_x):
        key = (K, want_x)
        b = self._bufs.get(key)
        if b is None:
            dev, n, f64 = self.device, self.n_envs, torch.float64
            b = {"obs": torch.empty(K, self._n_obs, n, dtype=f64, device=dev),
                 "reward": torch.empty(K, n, dtype=f64, device=dev),
                 "done": torch.empty(K, n, dtype=torch.uint8, device=dev),
                 "violated": torch.empty(K, n, dtype=torch.uint8, device=dev),
                 "x": torch.empty(K, self.Nx_oracle, n, dtype=f64, device=dev) if want_x else None}
            if len(self._bufs) > 4:
                self.host_sync()             # nothing cached may still be the source / target of a copy in flight
                self._bufs.clear()
            self._bufs[key] = b
        return b

    def _io(self, a, K, bufs, z=None) -> nat.IO:
        io = nat.IO()
        p = lambda t: t.data_ptr() if t is not None else None
        io.x, io.a, io.a_save = p(self._x), p(a), p(self._a_save)
        if self._sp_dev is not None:
            io.sp = p(self._sp_dev)
            nsp, n = len(self._sp_keys), self.n_envs
            io.sp_st, io.sp_sj, io.sp_se = (nsp * n, n, 1) if self._sp_per_env else (1, self.N, 0)
        if self._d_dev is not None:
            io.dist = p(self._d_dev)
            nd, n = len(self._d_keys), self.n_envs
            io.d_st, io.d_sj, io.d_se = (nd * n, n, 1) if self._d_per_env else (1, self.N, 0)
        io.z, io.tail = p(z), p(self._tail)
        io.x_traj, io.obs, io.reward = p(bufs["x"]), p(bufs["obs"]), p(bufs["reward"])
        io.done, io.violated = p(bufs["done"]), p(bufs["violated"])
        if self._cons is not None and not self._host_constraints:
            io.g0 = p(self._cons[0])
            io.g = self._cons[self.t + 1].data_ptr() if self.t + 1 < self.N else None
        io.done_state, io.status, io.h_carry, io.n_rhs = p(self._done_state), p(self._status), p(self._h_carry), p(self._n_rhs)
        io.u_prev = p(self._u_prev)
        io.theta = p(self._theta)
        return io

    # ------------------------------------------------------------------ reset (pcgym.py:263-349)
    def _begin_episode(self, episode=None):
        self.episode = int(episode) if episode is not None else self.episode + 1
        key = (self.seed + self.episode * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        self._desc.seed = key
        self._desc_step.seed = key

    # ------------------------------------------------------------------ fixed-step accuracy probe
    def fixed_step_error(self, n_probe: int = 512, intervals: int = 16, seed: int = 0) -> float:
        """Worst relative state difference between THIS env's fixed-step integrator and a tight adaptive run (Tsit5,
        rtol 1e-10) over a short probe rollout: `n_probe` envs from this env's own x0, uniform random actions over the
        whole action range, the first `intervals` control intervals, disturbances / parameters as configured.  A fixed
        step has no error estimate of its own (the headline scheme, Tsit5 x 1, holds the north-star 1e-6 at the
        documented CSTR interval dt = 0.25 with a 3x margin, but not at every dt), so reset() runs this once for
        fixed-step schemes and warns above 1e-6 (`integrator={"check": False}` turns it off).  Returns the error."""
        if self.integrator.get("adaptive", True) or self.integrator.get("scheme") == "rodas4":
            return 0.0
        n = int(min(n_probe, self.n_envs))
        K = int(max(1, min(intervals, self.N - 1)))

        def rows(v):                                      # per-env tables keep their first n rows
            a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            return a[:n] if (a.ndim == 2 and a.shape[0] == self.n_envs and self.n_envs > 1) else a
        # the probe compares states only: rewards and constraints (which may be Python callables rollout() cannot fuse) stay out
        base = {k: v for k, v in self.env_params.items()
                if k not in ("custom_reward", "integrator", "constraints", "done_on_cons_vio", "r_penalty", "cons_type")}
        base.update(n_envs=n, noise=False, device=str(self.device))
        if "x0" in base:
            base["x0"] = rows(base["x0"])
        if base.get("disturbances"):
            base["disturbances"] = {k: rows(v) for k, v in base["disturbances"].items()}
        if base.get("model_params"):
            base["model_params"] = {k: (rows(v) if np.ndim(v) else v) for k, v in base["model_params"].items()}
        g = torch.Generator(device="cpu").manual_seed(1234 + seed)
        nu = self._a_lo.shape[0]
        if self.normalise_a:
            a = torch.rand(K, nu, n, generator=g, dtype=torch.float64) * 2 - 1
        else:
            lo, hi = torch.as_tensor(self._a_lo)[None, :, None], torch.as_tensor(self._a_hi)[None, :, None]
            a = lo + (hi - lo) * torch.rand(K, nu, n, generator=g, dtype=torch.float64)
        a = a.to(self.device)
        xs = []
        for integ in (dict(self.integrator, check=False),
                      {"scheme": "tsit5", "adaptive": True, "rtol": 1e-10, "atol": 1e-12, "max_steps": 100000}):
            e = make_env(dict(base, integrator=integ))
            e.reset()
            xs.append(e.rollout(a)["x"].clone())
        ref = xs[1].abs().clamp_min(1e-300)
        err = ((xs[0] - xs[1]).abs() / ref)
        err = torch.where(torch.isfinite(err), err, torch.full_like(err, float("inf")))
        return float(err.max())

    def _check_fixed_step_once(self):
        if self._fixed_checked or self.integrator.get("adaptive", True) or not self.integrator.get("check", True):
            return
        self._fixed_checked = True
        try:
            err = self.fixed_step_error()
        except Exception as exc:  # noqa: BLE001  (advisory: a probe that cannot be built must not break reset())
            warnings.warn(f"pcgym_b200: fixed-step accuracy probe skipped ({type(exc).__name__}: {exc})", RuntimeWarning)
            return
        self.fixed_step_probe_error = err
        if not err <= 1e-6:
            it = self.integrator
            warnings.warn(
                f"pcgym_b200: integrator {it.get('scheme')} with {it.get('n_sub', 1)} fixed step(s) per control interval "
                f"(dt = {self.dt:g}) is {err:.1e} away from the converged solution on a probe rollout of this env -- the "
                "reference's integrator (and the 1e-6 parity bar) need more substeps, or the default adaptive scheme "
                "(integrator={'adaptive': True}); fixed steps carry no error estimate, so status stays 0",
                RuntimeWarning, stacklevel=3)

    def reset(self, seed: int = 0, **kwargs):
        self._ensure_device()
        self._join_pipe()
        self._check_fixed_step_once()
        self.t = 0
        self._begin_episode(kwargs.get("episode"))
        nx, n = self.Nx_oracle, self.n_envs
        x0 = self._x0_dev.clone()
        gen = torch.Generator(device=self.device)
        gen.manual_seed((self.seed * 1000003 + int(seed) + self.env_id0 + 7919 * self.episode) & (2 ** 63 - 1))
        if self.uncertainty_percentages is not None and "x0" in self.uncertainty_percentages:
            for idx, unc in enumerate(self.uncertainty_percentages["x0"]):         # iterates keys, as :286-288
                x0[idx] = self._perturb(x0[idx], float(unc), gen)
        self._x.copy_(x0[:nx])
        self._x_reset = x0[:nx]                                                    # rewind(to_x0=True) restores it
        parts = [x0]                                                               # [x | SP slots of x0]
        if self.disturbance_active and self._d_keys:
            d0 = self._d_dev[0] if self._d_per_env else self._d_dev[:, 0:1].expand(-1, n)
            parts.append(d0)
        if self.uncertainty:
            rows = []
            for name in self._unc_names:
                base = torch.full((n,), float(self.original_param_values[name]), dtype=torch.float64, device=self.device)
                if self.uncertainty_percentages is not None:
                    rows.append(self._perturb(base, float(self.uncertainty_percentages[name]), gen))
                else:
                    vals = torch.as_tensor(_np(self.empirical_distribution[name]), device=self.device)
                    rows.append(vals[torch.randint(len(vals), (n,), generator=gen, device=self.device)])
            self._tail.copy_(torch.stack(rows))
            parts.append(self._tail)
            if self.uncertainty_in_dynamics:         # the sampled values also drive the RHS of their env
                for j, name in enumerate(self._unc_names):
                    self._theta[self._theta_names.index(name)] = self._tail[j]
        if self.a_delta:
            self.a_save_init = _np(self.a_0).reshape(-1)
            self._a_save0 = torch.as_tensor(self.a_save_init, device=self.device)[:, None].expand(-1, n).contiguous()
            self._a_save.copy_(self._a_save0)
        self._state0 = torch.cat(parts, dim=0)                                     # [n_obs, n]
        self._done_state.zero_()
        self._episode_fresh = False
        self._status.zero_()
        self._h_carry.zero_()
        self._n_rhs.zero_()
        if self._cons is not None:
            self._cons.zero_()
        self.done = False
        obs = self._state0
        if self.normalise_o:
            obs = 2 * (obs - self._o_lo[:, None]) / (self._o_hi[:, None] - self._o_lo[:, None]) - 1
        else:
            obs = obs.clone()
        self.info["obs"] = obs.T.clone()
        if self.partial_observation:
            names = self.model.info()["states"]
            for i in range(nx):
                if names[i] not in self.partial_observation:
                    obs[i] = 0
        self._last_obs = obs
        self.info["r_init"] = 0
        if self._cons is not None:
            self.info["cons_info"] = self._cons.permute(1, 0, 2)
        self._numpy_mode = False
        return (self._single(obs) if self.n_envs == 1 else obs.T), self.info

    def _perturb(self, value, pct, gen):                                           # pcgym.py:255-261
        if self.distribution == "uniform":
            u = torch.rand(value.shape, generator=gen, device=self.device, dtype=torch.float64) * 2 - 1
            return value * (1 + u * pct)
        if self.distribution == "normal":
            return value + torch.randn(value.shape, generator=gen, device=self.device, dtype=torch.float64) * pct * value
        raise UnboundLocalError("distribution must be 'uniform' or 'normal'")

    @staticmethod
    def _single(t):
        return t[:, 0].detach().cpu().numpy().copy()

    # ------------------------------------------------------------------ step (pcgym.py:350-500)
    def step(self, action, z=None):
        if not self._dev_ready:
            raise RuntimeError("call reset() before step()")
        if self.t >= self.N - 1:
            raise IndexError(f"episode finished: t = {self.t}, N = {self.N} (list index out of range)")
        self._join_pipe()
        n, nu_a = self.n_envs, self._a_lo.shape[0]
        numpy_in = not isinstance(action, torch.Tensor)
        a = torch.as_tensor(np.asarray(action, dtype=np.float64)) if numpy_in else action
        a = a.to(device=self.device, dtype=torch.float64, non_blocking=True)
        if a.ndim <= 1:
            if n != 1 and a.numel() != n * nu_a:
                raise ValueError(f"action must have shape [{n}, {nu_a}]")
            a = a.reshape(n, nu_a)
        if a.shape != (n, nu_a):
            raise ValueError(f"action must have shape [{n}, {nu_a}], got {tuple(a.shape)}")
        self._a_step[0].copy_(a.T)
        if z is not None:                                    # externally supplied N(0,1) draws, [n_envs, nx]
            z = torch.as_tensor(np.asarray(z, dtype=np.float64)) if not isinstance(z, torch.Tensor) else z
            z = z.to(device=self.device, dtype=torch.float64).reshape(n, self.Nx_oracle).T.contiguous()[None]
        a_save_old = self._a_save.clone() if ((self._host_reward or self._host_constraints) and self.a_delta) else None
        t_old = self.t
        if self._host_constraints and t_old == 0:                                  # pcgym.py:416-420 (result discarded)
            self._constraint_check_host(self._uk_pre(self._a_step[0], a_save_old), numpy_in)
        out = self._launch(self._a_step, 1, want_x=False, z=z, desc=self._desc_step)
        obs_soa, r, done, violated = out["obs"][0], out["reward"][0], out["done"][0], out["violated"][0]
        if self._copy_outputs or self._host_reward or self._host_constraints:
            obs_soa, r = obs_soa.clone(), r.clone()
        if self._host_constraints:                                                 # pcgym.py:444-446, 556-557, 529-530
            violated = self._constraint_check_host(self._uk(self._a_step[0], a_save_old, t_old), numpy_in).view(torch.uint8)
            vb = violated.view(torch.bool)
            if self.done_on_constraint:
                done = done.clone()
                done |= violated
                self._done_state |= violated
            if self.r_penalty and not self._host_reward:
                if self.reward == "SP_reward_fn":
                    r = r - 1000.0 * len(self._sp_keys) * vb                        # once per SP key
                elif self.t == self.N - 1:
                    r = r - 1000.0 * vb
        if self._host_reward:
            r, obs_soa = self._custom_reward(obs_soa, self._a_step[0], a_save_old, t_old, violated.view(torch.bool), numpy_in)
        elif self._obs_mask_bits:
            # info["obs"] keeps the unmasked normalised observation (pcgym.py:489), the returned one is masked (:495-498)
            self.info["obs"] = obs_soa.T.clone()
            obs_soa = obs_soa.clone() if not self._copy_outputs else obs_soa
            obs_soa[self._mask_rows] = 0
        self._last_obs = obs_soa
        if not self._obs_mask_bits or self._host_reward:
            self.info["obs"] = obs_soa.T
        self.info["violated"] = violated.view(torch.bool).clone() if self._copy_outputs else violated.view(torch.bool)
        self.info["status"] = self._status
        if n == 1 and numpy_in:
            self.done = bool(done.item())
            return self._single(obs_soa), float(r.item()) if isinstance(r, torch.Tensor) else float(r), self.done, False, self.info
        # the kernel writes 0 / 1 bytes: reinterpreting them as bool costs no launch
        done_b = done.view(torch.bool).clone() if self._copy_outputs else done.view(torch.bool)
        self.done = done_b
        if self._trunc is None:
            self._trunc = torch.zeros(n, dtype=torch.bool, device=self.device)
        return obs_soa.T, r, done_b, self._trunc, self.info

    def _uk(self, a_soa, a_save_old, t_old):
        """Physical input vector uk [Nu, n] of the interval just integrated (pcgym.py:372-414), for custom rewards."""
        lo = torch.as_tensor(self._a_lo, device=self.device)[:, None]
        hi = torch.as_tensor(self._a_hi, device=self.device)[:, None]
        u = a_soa
        if self.normalise_a is True:
            u = (u + 1) * (hi - lo) / 2 + lo
        if self.normalise_a and self.a_delta:
            u = a_save_old + ((u + 1) * (hi - lo) / 2 + lo)
        rows = [u]
        if self.disturbance_active:
            params = self.model.info()["parameters"]
            for k in (self.model.info().get("disturbances") or []):
                if k in self._d_keys:
                    j = self._d_keys.index(k)
                    v = self._d_dev[t_old + 1, j] if self._d_per_env else self._d_dev[j, t_old + 1].expand(self.n_envs)
                else:
                    v = torch.full((self.n_envs,), float(params[str(k)]), dtype=torch.float64, device=self.device)
                rows.append(v[None])
        return torch.cat(rows, dim=0)

    def _uk_pre(self, a_soa, a_save_old):
        """uk as the t == 0 pre-check sees it (pcgym.py:416-420): this interval's input, disturbances of t + 1."""
        return self._uk(a_soa, a_save_old, 0)

    def _constraint_check_host(self, uk, numpy_in):
        """constraint_check / con_checker (pcgym.py:560-615) for a callable that could not be fused: evaluated on the
        physical state vector and uk -- after the reference's un-normalisation of already-physical values when
        constraint_units == "reference" (pcgym.py:597-608).  n_envs == 1: numpy vectors like the reference;
        batched: numpy arrays [Nx, n] and [Nu, n] (state-major, so g written for one env -- x[1], u[0] -- vectorises
        as is) returning [n_con, n].  Writes cons_info[:, t, :], returns violated [n] (bool tensor)."""
        st = self.state
        st = st.T if isinstance(st, torch.Tensor) else torch.as_tensor(st, device=self.device)[:, None]     # [Nx, n]
        st, uk = st.detach().cpu().numpy().astype(np.float64), uk.detach().cpu().numpy().astype(np.float64)
        if self.env_params.get("constraint_units", "reference") == "reference":
            if self.normalise_a is True:
                lo, hi = self._a_lo[:, None], self._a_hi[:, None]
                uk = (uk + 1) * (hi - lo) / 2 + lo
            if self.normalise_o is True:
                lo = self.observation_space_base.low.astype(np.float64)[:, None]
                hi = self.observation_space_base.high.astype(np.float64)[:, None]
                st = (st + 1) * (hi - lo) / 2 + lo
        if self.n_envs == 1:
            g = np.asarray(self.constraints(st[:, 0], uk[:, 0]), dtype=np.float64).reshape(self.n_con, 1)
        else:
            g = np.asarray(self.constraints(st, uk), dtype=np.float64).reshape(self.n_con, self.n_envs)
        gd = torch.from_numpy(g).to(self.device)
        self._cons[self.t] = gd
        return (gd > 0).any(dim=0)

    def _custom_reward(self, obs_phys, a_soa, a_save_old, t_old, violated, numpy_in):
        """Python-callable custom reward (pcgym.py:470-471): `f(env, obs, uk, violated)` with the NOISY
        physical observation.  A callable cannot be fused, so the kernel ran with reward NONE and emitted
        physical observations; the callable is evaluated here and the observation is normalised / masked
        afterwards with torch ops.  n_envs == 1 with numpy input calls it exactly like the reference (numpy
        vectors, Python bool); batched envs pass torch tensors [n, Nx], [n, Nu], [n]."""
        uk = self._uk(a_soa, a_save_old, t_old)
        if self.n_envs == 1 and numpy_in:
            r = self.custom_reward_f(self, self._single(obs_phys), self._single(uk), bool(violated.item()))
            r = float(np.asarray(r).reshape(-1)[0])
        else:
            r = self.custom_reward_f(self, obs_phys.T, uk.T, violated)
            r = torch.as_tensor(r, dtype=torch.float64, device=self.device).reshape(self.n_envs)
        obs = obs_phys
        if self.normalise_o is True:
            obs = 2 * (obs - self._o_lo[:, None]) / (self._o_hi[:, None] - self._o_lo[:, None]) - 1
        for i in range(self.Nx_oracle):
            if self._obs_mask_bits >> i & 1:
                obs[i] = 0
        return r, obs

    # ------------------------------------------------------------------ rollout: K intervals, one launch
    def rollout(self, actions, z=None, want_x=True, copy=False):
        """actions: [K, nu_a, n_envs] (struct-of-arrays, preferred) or [K, n_envs, nu_a]; device or host tensor.
        Returns SoA tensors: obs [K, Nx, n], reward [K, n], done/violated [K, n] (uint8), x [K, nx, n].
        The returned tensors are the env's own output buffers for this (K, want_x): the next rollout() /
        rollout_host() of the same shape OVERWRITES them (no allocation on the hot path).  Pass copy=True (or clone
        what you keep) to own the result."""
        if not self._dev_ready:
            raise RuntimeError("call reset() before rollout()")
        self._join_pipe()
        n, nu_a = self.n_envs, self._a_lo.shape[0]
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions, dtype=np.float64))
        if a.ndim != 3:
            raise ValueError("actions must be [K, nu_a, n_envs]")
        if a.shape[1:] != (nu_a, n):
            if a.shape[1:] == (n, nu_a):
                a = a.permute(0, 2, 1)
            else:
                raise ValueError(f"actions must be [K, {nu_a}, {n}] (or [K, {n}, {nu_a}]), got {tuple(a.shape)}")
        if self._host_reward:
            raise NotImplementedError("rollout() needs a fused reward: use SP / batch reward or custom_reward="
                                      "{'kind': 'ocp_tracking', ...}; Python-callable rewards work through step()")
        if self._host_constraints:
            raise NotImplementedError("rollout() needs fusable (affine or box) constraints; a non-affine callable is "
                                      "evaluated unfused through step()")
        a = a.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()
        K = a.shape[0]
        if self.t + K > self.N - 1:
            raise IndexError(f"rollout of {K} steps from t = {self.t} exceeds the episode (N - 1 = {self.N - 1} steps)")
        if z is not None:
            z = z.to(device=self.device, dtype=torch.float64).contiguous()
            if z.shape != (K, self.Nx_oracle, n):
                raise ValueError(f"z must be [K, {self.Nx_oracle}, {n}]")
        out = self._launch(a, K, want_x=want_x, z=z)
        return {k: (v.clone() if v is not None else None) for k, v in out.items()} if copy else out

    HOST_OUTPUTS = ("obs", "reward", "done", "violated", "x", "obs_last", "return")

    def rollout_host(self, actions, out=None, chunks: int = 8, want_x: bool = False, wait: bool = True, outputs=None):
        """`rollout` for HOST buffers: `actions` [K, nu_a, n_envs] float64 in (ideally pinned) host memory; returns
        pinned host tensors, reusing `out` when given.  `outputs` names what is brought back (default
        ("obs", "reward", "done"), plus "x" with `want_x`):
            per interval   "obs" [K, Nx, n]  "reward" [K, n]  "done" [K, n]  "violated" [K, n]  "x" [K, nx, n]
            reduced        "obs_last" [Nx, n] (observation after the last interval)   "return" [n] (sum of the K rewards,
                           reduced on the device)
        A consumer that needs the returns and the final observation of an episode moves 32 B per env instead of
        3.3 KB.  The K intervals are cut into `chunks` time slices -- contiguous in both the action and the output
        layouts -- and pipelined over three streams: the host->device copy of slice c+1 and the device->host copy of
        slice c-1 overlap the kernel of slice c (PCIe is full duplex), so the call costs about max(H2D, D2H) instead
        of their sum.  The result is complete when the call returns (it synchronises).  With `wait=False` the call
        only enqueues: successive calls then overlap as well (slice j of the next call waits for the device->host
        copy of slice j of this one, nothing more), and `host_sync()` must be called before `out` is read --
        alternate two `out` sets to keep one readable."""
        if not self._dev_ready:
            raise RuntimeError("call reset() before rollout_host()")
        if self._host_reward or self._host_constraints:
            raise NotImplementedError("rollout_host() needs a fused reward and fusable constraints (see rollout())")
        if outputs is None:
            outputs = ("obs", "reward", "done") + (("x",) if want_x else ())
        outputs = tuple(outputs)
        bad = [k for k in outputs if k not in self.HOST_OUTPUTS]
        if bad:
            raise ValueError(f"unknown outputs {bad}; choose from {self.HOST_OUTPUTS}")
        n, nu_a = self.n_envs, self._a_lo.shape[0]
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions, dtype=np.float64))
        if a.device.type != "cpu" or a.dtype != torch.float64 or a.ndim != 3 or a.shape[1:] != (nu_a, n):
            raise ValueError(f"actions must be a float64 host tensor [K, {nu_a}, {n}]")
        a = a.contiguous()
        K = a.shape[0]
        if self.t + K > self.N - 1:
            raise IndexError(f"rollout of {K} steps from t = {self.t} exceeds the episode (N - 1 = {self.N - 1} steps)")
        if self._desc.noise == nat.NOISE["external"]:
            raise ValueError("noise_mode='external' is not supported by rollout_host()")
        want_x = want_x or "x" in outputs
        dev = self.device
        keys = tuple(k for k in outputs if k in ("obs", "reward", "done", "violated", "x"))
        bufs = self._out_bufs(K, want_x)
        if out is None:
            out = {k: torch.empty(bufs[k].shape, dtype=bufs[k].dtype).pin_memory() for k in keys}
            if "obs_last" in outputs:
                out["obs_last"] = torch.empty(bufs["obs"].shape[1:], dtype=torch.float64).pin_memory()
            if "return" in outputs:
                out["return"] = torch.empty(n, dtype=torch.float64).pin_memory()
        if self._pipe is None:
            self._pipe = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        s_in, s_out = self._pipe
        main = torch.cuda.current_stream(dev)
        a_dev = self._bufs.get(("a_dev", K))
        if a_dev is None:
            a_dev = self._bufs[("a_dev", K)] = torch.empty(K, nu_a, n, dtype=torch.float64, device=dev)
        edges = [int(e) for e in np.linspace(0, K, max(1, min(int(chunks), K)) + 1).astype(int)]
        sig = (K, tuple(edges), want_x, outputs)
        prev = self._pipe_prev if self._pipe_prev is not None and self._pipe_prev[0] == sig else None
        if prev is None:                          # different slicing (or first call): order behind everything in flight
            main.wait_stream(s_out)
            s_in.wait_stream(main)
            s_out.wait_stream(main)
        ev_ks, ev_outs = [], []
        last = len(edges) - 2
        for j, (k0, k1) in enumerate(zip(edges[:-1], edges[1:])):
            with torch.cuda.stream(s_in):
                if prev is not None:
                    s_in.wait_event(prev[1][j])   # the previous call's kernel has finished reading this action slice
                a_dev[k0:k1].copy_(a[k0:k1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            main.wait_event(ev_in)
            if prev is not None:
                main.wait_event(prev[2][j])       # ... and its results have left this output slice
            self._launch(a_dev[k0:k1], k1 - k0, want_x, bufs={k: (v[k0:k1] if v is not None else None) for k, v in bufs.items()})
            ret_dev = None
            if j == last and "return" in outputs:
                ret_dev = self._bufs.get(("ret", K))
                if ret_dev is None:
                    ret_dev = self._bufs[("ret", K)] = torch.empty(n, dtype=torch.float64, device=dev)
                torch.sum(bufs["reward"], dim=0, out=ret_dev)
            ev_k = torch.cuda.Event()
            ev_k.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_k)
                for k in keys:
                    out[k][k0:k1].copy_(bufs[k][k0:k1], non_blocking=True)
                if j == last:
                    if "obs_last" in outputs:
                        out["obs_last"].copy_(bufs["obs"][K - 1], non_blocking=True)
                    if ret_dev is not None:
                        out["return"].copy_(ret_dev, non_blocking=True)
                ev_o = torch.cuda.Event()
                ev_o.record(s_out)
            ev_ks.append(ev_k)
            ev_outs.append(ev_o)
        self._pipe_prev = (sig, ev_ks, ev_outs)
        if wait:
            self.host_sync()
        return out

    def _join_pipe(self):
        """Order the caller's stream behind copies still in flight from rollout_host(wait=False)."""
        if self._pipe_prev is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._pipe[1])
            self._pipe_prev = None

    def host_sync(self):
        """Complete every `rollout_host(..., wait=False)` issued so far (results are in their `out` buffers)."""
        if self._pipe is not None:
            main = torch.cuda.current_stream(self.device)
            main.wait_stream(self._pipe[1])
            main.synchronize()
        self._pipe_prev = None

    def _launch(self, a, K, want_x, z=None, bufs=None, desc=None):
        d = self._desc if desc is None else desc
        if d.noise == nat.NOISE["external"] and z is None:
            raise ValueError("noise_mode='external' needs the standard normals `z`")
        if bufs is None:
            bufs = self._out_bufs(K, want_x)
        io = self._io(a, K, bufs, z)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        flags0 = d.flags
        if self._episode_fresh:              # first launch after rewind(): the sticky done flag is not read (no memset needed)
            d.flags = flags0 | nat.F_EPISODE_START
        try:
            with torch.cuda.device(self.device):
                nat.check(nat.lib().pcg_rollout(ctypes.byref(d), ctypes.byref(io), self.n_envs, self.t, K,
                                                ctypes.c_void_p(stream)))
        finally:
            d.flags = flags0
        self._episode_fresh = False
        self.t += K
        if d.ocp_mode & 3 == 2:              # io.u_prev is valid from now on (it survives reset(), as `self.u_prev` does)
            self._desc.ocp_mode = self._desc_step.ocp_mode = (d.ocp_mode & ~3) | 1
        return bufs

    def rewind(self, t: int = 0, episode=None, to_x0: bool = False):
        """Begin a new episode from the CURRENT device state: only the time index and the sticky done flag
        are cleared (no x0 copy, no allocation) and the noise key moves on to the next episode (or to `episode`).
        Continuing-task rollouts and the headline benchmark use this.  `to_x0=True` also puts back the model state
        (and the a_delta accumulator) the last reset() produced -- one device-to-device copy, no host work and no new
        random draws: what a benchmark pass over a model that does not survive arbitrary action sequences for long
        (crystallization) starts from.  `status` is NOT cleared (only reset() does), so trouble in any pass stays visible."""
        if not self._dev_ready:
            raise RuntimeError("call reset() before rewind()")
        self._begin_episode(episode)
        self.t = int(t)
        self._episode_fresh = True           # instead of a memset of the sticky done flag: PCG_F_EPISODE_START on the next launch
        self.done = False
        if to_x0:
            self._x.copy_(self._x_reset)
            if self.a_delta:
                self._a_save.copy_(self._a_save0)

    # ------------------------------------------------------------------ reference attributes
    @property
    def obs(self):
        return self._last_obs.T if self.n_envs > 1 else self._single(self._last_obs)

    @property
    def state(self):
        """Physical state vector [x | SP[t-1] | disturbances(t) | uncertain params] (pcgym.py:284-316, 409-438)."""
        if self.t == 0:
            s = self._state0.clone()
        else:
            parts = [self._x]
            n = self.n_envs
            if self._sp_dev is not None and not self._sp_hidden:
                parts.append(self._sp_dev[self.t - 1] if self._sp_per_env else self._sp_dev[:, self.t - 1: self.t].expand(-1, n))
            if self._d_dev is not None:
                parts.append(self._d_dev[self.t] if self._d_per_env else self._d_dev[:, self.t: self.t + 1].expand(-1, n))
            if self._tail is not None:
                parts.append(self._tail)
            s = torch.cat(parts, dim=0)
        return s.T if self.n_envs > 1 else self._single(s)

    @property
    def a_save(self):
        if not self.a_delta:
            raise AttributeError("a_save")
        if not self._dev_ready:
            return self.a_0
        return self._a_save.T if self.n_envs > 1 else self._single(self._a_save)

    @property
    def status(self):
        return self._status

    @property
    def rhs_evaluations(self):
        return self._n_rhs

    # ------------------------------------------------------------------ pcgym.py:617-692
    def get_rollouts(self, policies: dict, reps: int, oracle: bool = False, dist_reward: bool = False,
                     MPC_params=False, cons_viol: bool = False, **kw):
        """Closed-loop rollouts of `policies`, `reps` repetitions each (the repetitions are the env batch of one
        `policy_eval`); returns (evaluator, data) like the reference."""
        from .policy_evaluation import policy_eval
        evaluator = policy_eval(make_env, policies, reps, self.env_params, oracle, MPC_params, **kw)
        return evaluator, evaluator.get_rollouts()

    def plot_rollout(self, policies: dict, reps: int, oracle: bool = False, dist_reward: bool = False,
                     MPC_params=False, cons_viol: bool = False, save_fig: bool = False, **kw):
        from .policy_evaluation import policy_eval
        evaluator = policy_eval(make_env, policies, reps, self.env_params, oracle, MPC_params, cons_viol, save_fig, **kw)
        data = evaluator.get_rollouts()
        evaluator.plot_data(data, dist_reward)              # raises: plotting is out of scope
        return evaluator, data

    def close(self):
        self._bufs = {}
