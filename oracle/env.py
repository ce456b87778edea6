"""Oracle environment: single-env numpy restatement of `pcgym.make_env`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the reference order of
operations exactly, quirks included; citations are to
`/root/reference/src/pcgym/pcgym.py`.  The ODE advance is delegated to
`oracle.integrate` (converged scipy integrator instead of CasADi/CVODES).

Differences from the reference that are deliberate and visible:
  * `normal` hook: the Gaussian draws of the measurement-noise block come from
    `normal(size)` (default `np.random.normal(0, 1, size)`, the reference's
    source) so tests can replay recorded draws.
  * `integrator` selects "truth" (DOP853 1e-12), "lsoda" (CVODES-default
    stand-in, for timing) or "dense" (RK4 x n_sub).
  * `env_params["constraint_units"] = "physical"` (the same opt-in the CUDA
    env offers) skips constraint_check's un-normalisation of values that are
    already physical (pcgym.py:597-608); the default replicates it.
"""
from __future__ import annotations

import copy

import numpy as np

from . import integrate as _int
from . import rhs as _rhs


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (gymnasium is not in this image)."""

    def __init__(self, low, high, dtype=np.float64):
        self.low = np.asarray(low, dtype=np.float64)
        self.high = np.asarray(high, dtype=np.float64)
        self.shape = self.low.shape
        self.dtype = dtype

    def sample(self):
        return np.random.uniform(self.low, self.high)


class OracleModel:
    """Registry model with the reference's duck-typed surface: __call__(x,u), info()."""

    def __init__(self, name, int_method="casadi"):
        meta = _rhs.MODELS[name]
        self.name = name
        self.int_method = int_method
        self.states = list(meta["states"])
        self.inputs = list(meta["inputs"])
        self.disturbances = list(meta["disturbances"])
        self.parameters = dict(meta["params"])

    def __getattr__(self, k):                     # getattr(model, "k0") like the dataclasses
        params = self.__dict__.get("parameters", {})
        if k in params:
            return params[k]
        raise AttributeError(k)

    def __call__(self, x, u):
        return _rhs.RHS[self.name](x, u, self.parameters)

    def info(self):
        return {"parameters": dict(self.parameters), "states": self.states, "inputs": self.inputs,
                "disturbances": self.disturbances}


class OracleEnv:
    def __init__(self, env_params, integrator="truth", normal=None, n_sub=64):
        if not isinstance(env_params, dict):
            raise ValueError("env_params must be a dictionary")          # pcgym.py:40-41
        P = self.env_params = copy.deepcopy(env_params)
        self._integrator = integrator
        self._n_sub = n_sub
        self._normal = normal if normal is not None else (lambda size: np.random.normal(0, 1, size))

        # pcgym.py:56-61
        self.a_delta = P.get("a_delta", False)
        if self.a_delta:
            self.a_0 = P["a_0"]
        self.normalise_a = P.get("normalise_a", True)
        self.normalise_o = P.get("normalise_o", True)

        # pcgym.py:68-92
        a_lo, a_hi = np.asarray(P["a_space"]["low"], float), np.asarray(P["a_space"]["high"], float)
        n_a = a_lo.shape[0]
        self.action_space = Box(-np.ones(n_a), np.ones(n_a)) if self.normalise_a else Box(a_lo, a_hi)
        self.observation_space_base = Box(P["o_space"]["low"], P["o_space"]["high"])

        # pcgym.py:94-111
        self.maximise_reward = True
        self.SP = P.get("SP")
        self.custom_reward = P.get("custom_reward") is not None
        if self.custom_reward:
            self.custom_reward_f = P["custom_reward"]
            self.reward = None
        elif self.SP is not None:
            self.reward = "SP_reward_fn"
        else:
            self.reward = "batch_reward_fn"
            self.reward_states = P["reward_states"]
            self.maximise_reward = P["maximise_reward"]
        self.N = P["N"]
        self.tsim = P["tsim"]
        self.x0 = P["x0"]
        self.integration_method = P.get("integration_method", "casadi")
        self.dt = self.tsim / self.N
        self.done = False

        # pcgym.py:113-125
        self.constraint_active = False
        self.r_penalty = False
        self.info = {}
        if P.get("constraints") is not None:
            self.constraints = P["constraints"]
            self.done_on_constraint = P["done_on_cons_vio"]
            self.r_penalty = P["r_penalty"]
            self.constraint_active = True
            self.n_con = np.asarray(self.constraints(self.x0, self.action_space.sample())).shape[0]
            self.info["cons_info"] = np.zeros((self.n_con, self.N, 1))

        # pcgym.py:127-158
        if P.get("custom_model") is not None:
            self.model = P["custom_model"]
            self.model.int_method = self.integration_method
        else:
            name = P.get("model")
            if name not in _rhs.MODELS:
                raise ValueError(f"Model '{name}' not found in model_mapping.")
            self.model = OracleModel(name, self.integration_method)

        # pcgym.py:160-165
        minfo = self.model.info()
        self.Nx_oracle = len(minfo["states"])
        self.Nx = self.Nx_oracle + (len(self.SP) if self.SP is not None else 0)
        self.Nu = len(minfo["inputs"])

        # pcgym.py:167-199
        self.disturbance_active = False
        self.Nd = self.Nd_model = 0
        if P.get("disturbances") is not None:
            self.disturbance_active = True
            self.disturbances = P["disturbances"]
            self.Nd = len(self.disturbances)
            self.Nd_model = len(minfo["disturbances"])
            self.Nu += self.Nd_model
            self.Nx += self.Nd
            self._extend_obs(P["disturbance_bounds"])

        # pcgym.py:212-252
        self.uncertainty = False
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
            self._extend_obs(P["uncertainty_bounds"])

        if self.normalise_o:
            n = self.observation_space_base.low.shape[0]
            self.observation_space = Box(-np.ones(n), np.ones(n))
        else:
            self.observation_space = self.observation_space_base

        # pcgym.py:63-66, 208-211
        self.noise_percentage = P.get("noise_percentage")
        if self.noise_percentage is not None:
            self.noise_percentage_float = isinstance(self.noise_percentage, float)
        self.partial_observation = P.get("partial_observation") or False

    # -- helpers -----------------------------------------------------------
    def _extend_obs(self, bounds):
        lo = np.concatenate((self.observation_space_base.low, np.asarray(bounds["low"], float)))
        hi = np.concatenate((self.observation_space_base.high, np.asarray(bounds["high"], float)))
        self.observation_space_base = Box(lo, hi)

    def _perturb(self, value, pct, distribution):                      # pcgym.py:255-261
        if distribution == "uniform":
            return value * (1 + np.random.uniform(-pct, pct))
        if distribution == "normal":
            return np.random.normal(value, pct * value)
        raise UnboundLocalError("noisy_value")                         # reference falls through

    def _emit(self, obs):
        """normalise (pcgym.py:332-342 / 483-493) + partial-observation mask (:344-347 / 495-498)."""
        if self.normalise_o:
            lo, hi = self.observation_space_base.low, self.observation_space_base.high
            out = 2 * (obs - lo) / (hi - lo) - 1
        else:
            out = obs
        self.info["obs"] = copy.deepcopy(out)
        if self.partial_observation:
            names = self.model.info()["states"]
            for i in range(self.Nx_oracle):
                if names[i] not in self.partial_observation:
                    out[i] = 0
        return out

    def _advance(self, x, uk):                                          # integrator.py:90-107
        if self._integrator == "truth":
            return _int.advance_truth(self.model, x, uk, self.dt)
        if self._integrator == "lsoda":
            return _int.advance_lsoda(self.model, x, uk, self.dt)
        if self._integrator == "dense":
            return _int.advance_dense(self.model, x, uk, self.dt, n_sub=self._n_sub)
        raise ValueError(self._integrator)

    # -- reset (pcgym.py:263-349) -------------------------------------------
    def reset(self, seed=0, **kwargs):
        self.t = 0
        state = np.array(copy.deepcopy(self.env_params["x0"]), dtype=np.float64)
        if self.uncertainty_percentages is not None and "x0" in self.uncertainty_percentages:
            for idx, unc in enumerate(self.uncertainty_percentages["x0"]):      # :286-288 (iterates keys)
                state[idx] = self._perturb(state[idx], unc, self.distribution)
        if self.disturbance_active:
            first = [self.disturbances[k][0] for k in self.model.info()["disturbances"] if k in self.disturbances]
            state = np.concatenate((state, first))
        if self.uncertainty:
            drawn = []
            if self.uncertainty_percentages is not None:
                for name, pct in self.uncertainty_percentages.items():
                    if name != "x0":
                        v = self._perturb(self.original_param_values[name], pct, self.distribution)
                        drawn.append(v)          # observation only: never reaches the dynamics (:281 vs :308)
            else:
                for name in self.empirical_distribution:
                    drawn.append(np.random.choice(self.empirical_distribution[name]))
            state = np.concatenate((state, drawn))
        if self.a_delta:
            self.a_save = self.a_0
        self.state = state
        self.obs = copy.deepcopy(state)
        self.done = False
        out = self._emit(self.obs)
        self.info["r_init"] = 0
        return out, self.info

    # -- step (pcgym.py:350-500) -------------------------------------------
    def step(self, action):
        P = self.env_params
        a_lo, a_hi = P["a_space"]["low"], P["a_space"]["high"]
        if self.normalise_a is True:                                            # :372-375
            action = (action + 1) * (a_hi - a_lo) / 2 + a_lo
        if self.normalise_a and self.a_delta:                                   # :376-383
            action = (action + 1) * (a_hi - a_lo) / 2 + a_lo
            action = self.a_save + action
            self.a_save = np.clip(action, P["a_space_act"]["low"], P["a_space_act"]["high"])

        if self.disturbance_active:                                             # :386-412
            minfo = self.model.info()
            names = minfo["disturbances"]
            uk = np.zeros(self.Nu)
            uk[: self.Nu - len(names)] = action
            given = []
            for i, k in enumerate(names):
                if k in self.disturbances:
                    uk[self.Nu - self.Nd_model + i] = self.disturbances[k][self.t + 1]
                    given.append(self.disturbances[k][self.t + 1])
                else:
                    uk[self.Nu - self.Nd_model + i] = minfo["parameters"][str(k)]
            off = self.Nx_oracle + len(self.SP)
            if self.uncertainty_percentages is not None:
                off += len(self.uncertainty_percentages)
            self.state[off:] = given
        else:
            uk = action                                                         # :414

        if self.t == 0 and self.constraint_active:                              # :416-420
            self.constraint_check(self.state, uk)

        self.state[: self.Nx_oracle] = self._advance(self.state[: self.Nx_oracle], np.asarray(uk, float))

        if self.SP is not None:                                                 # :432-438 (t before increment)
            self.state[self.Nx_oracle: self.Nx_oracle + len(self.SP)] = np.array(
                [self.SP[k][self.t] for k in self.SP])
        self.t += 1                                                             # :441

        violated = False
        if self.constraint_active:                                              # :444-446
            violated = self.constraint_check(self.state, uk)
        if self.t == self.N - 1:                                                # :448-449
            self.done = True

        self.obs = copy.deepcopy(self.state)                                    # :453-466
        if P.get("noise", False):
            if self.noise_percentage_float:
                pct = P.get("noise_percentage", 0)
                self.obs[: self.Nx_oracle] += self._normal(self.Nx_oracle) * self.state[: self.Nx_oracle] * pct
            else:
                names = self.model.info()["states"]
                for i in range(self.Nx_oracle):
                    if names[i] in self.noise_percentage:
                        self.obs[i] += (self._normal(1) * self.state[i] * self.noise_percentage[str(names[i])])[0]

        if self.custom_reward:                                                  # :470-482
            rew = self.custom_reward_f(self, self.obs, uk, violated)
        elif self.reward == "SP_reward_fn":
            rew = self.SP_reward_fn(self.state, violated)
        else:
            rew = self.batch_reward_fn(self.state, violated)

        out = self._emit(self.obs)
        return out, rew, self.done, False, self.info

    # -- rewards (pcgym.py:502-558) -------------------------------------------
    def SP_reward_fn(self, state, c_violated):
        r = 0.0
        names = self.model.info()["states"]
        scale = self.env_params.get("r_scale", {})
        for k in self.SP:
            i = names.index(k)
            r += (-((state[i] - np.array(self.SP[k][self.t])) ** 2)) * scale.get(k, 1)
            if self.r_penalty and c_violated:                                   # once per SP key (:556-557)
                r -= 1000
        return r

    def batch_reward_fn(self, state, c_violated):
        r = 0.0
        if self.t == self.N - 1:
            names = self.model.info()["states"]
            scale = self.env_params.get("r_scale", {})
            for s in self.reward_states:
                if str(s) in names:
                    v = state[names.index(s)] * scale.get(s, 1)
                    r = r + v if self.maximise_reward else r - v
            if self.r_penalty and c_violated:
                r -= 1000
        return r

    # -- constraints (pcgym.py:560-615) ---------------------------------------
    def constraint_check(self, state, inputs):
        P = self.env_params
        quirk = P.get("constraint_units", "reference") == "reference"
        if self.normalise_a is True and quirk:   # un-normalises values that are already physical (:597-600)
            inputs = (inputs + 1) * (P["a_space"]["high"] - P["a_space"]["low"]) / 2 + P["a_space"]["low"]
        if self.normalise_o is True and quirk:                                  # :602-608
            lo, hi = self.observation_space_base.low, self.observation_space_base.high
            state = (state + 1) * (hi - lo) / 2 + lo
        g = np.asarray(self.constraints(state, inputs), dtype=np.float64)
        self.info["cons_info"][:, self.t, :] = g.reshape(g.shape[0], 1)
        violated = bool(np.any(self.info["cons_info"][:, self.t, :] > 0))
        if violated and self.done_on_constraint:
            self.done = True
        return violated
