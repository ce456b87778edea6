"""Shared env_params for the step-semantics parity cases.

Used by `tests/golden/make_golden.py` (which runs them through the reference's
own `make_env`, in the build container) and by the oracle / CUDA parity tests.
Configurations follow the reference's README, docs, tests and paper scripts;
sources are cited per case.  Pure numpy: no reference import here.
"""
from __future__ import annotations

import numpy as np


def _cstr_base(N=40):
    # README.md:20-54 quick-start, tsim scaled so dt = 0.25
    return {
        "N": N, "tsim": N * 0.25,
        "SP": {"Ca": [0.85] * (N // 2) + [0.9] * (N - N // 2)},
        "o_space": {"low": np.array([0.7, 300, 0.8]), "high": np.array([1, 350, 0.9])},
        "a_space": {"low": np.array([295.0]), "high": np.array([302.0])},
        "x0": np.array([0.8, 330, 0.8]),
        "model": "cstr",
        "r_scale": {"Ca": 1e3},
    }


def cstr_basic():
    """README quick-start + float noise (tests/environment/test_make_env_basic.py:18-20)."""
    p = _cstr_base()
    p.update(noise=True, noise_percentage=0.001)
    return p


def cstr_dist_cons():
    """Disturbances (example_notebooks/Disturbances.ipynb) + constraints (docs/guides/constraints.md:38),
    normalisation ON, so constraint_check's double un-normalisation is live (pcgym.py:597-608)."""
    p = _cstr_base()
    N = p["N"]
    rng = np.random.default_rng(11)
    p.update(
        disturbances={"Ti": np.repeat(rng.uniform(340, 360, N // 4), 4), "Caf": rng.uniform(0.9, 1.1, N)},
        disturbance_bounds={"low": np.array([330.0, 0.8]), "high": np.array([370.0, 1.2])},
        constraints=lambda x, u: np.array([319 - x[1], x[1] - 331]).reshape(-1,),
        done_on_cons_vio=False, r_penalty=True,
    )
    return p


def cstr_nonorm_cons():
    """Physical-unit actions/observations, constraint on T, episode ends on violation."""
    p = _cstr_base()
    p.update(
        normalise_a=False, normalise_o=False,
        constraints=lambda x, u: np.array([322.5 - x[1], x[1] - 331, u[0] - 301.5]).reshape(-1,),
        done_on_cons_vio=True, r_penalty=True,
        noise=True, noise_percentage={"T": 0.002},
    )
    return p


def cstr_one_dist():
    """Only Caf supplied: Ti falls back to the model default (pcgym.py:399-405); Nd=1, Nd_model=2."""
    p = _cstr_base()
    N = p["N"]
    p.update(
        disturbances={"Caf": np.random.default_rng(5).uniform(0.9, 1.1, N)},
        disturbance_bounds={"low": np.array([0.8]), "high": np.array([1.2])},
        partial_observation=["Ca"],
    )
    return p


def cstr_batch_reward():
    """No SP: terminal batch reward (pcgym.py:502-532)."""
    p = _cstr_base(N=12)
    del p["SP"]
    p.update(reward_states=["Ca", "T"], maximise_reward=False, r_scale={"Ca": 10.0},
             o_space={"low": np.array([0.7, 300]), "high": np.array([1, 350])}, x0=np.array([0.8, 330.0]))
    return p


def ocp_reward(R=0.1, delta_u=True):
    """The "OCP tracking" reward every paper script passes as custom_reward (same mathematics as
    pc-gym_paper/train_policies/cstr/custom_reward.py:3-39, written independently): squared setpoint error
    of the o_space-normalised noisy observation, weighted by r_scale, plus R * (change of the a_space-normalised
    input)^2; negative cost.  Keeps `u_prev` on the env object like the original."""
    def reward(env, x, u, violated):
        P = env.env_params
        names = env.model.info()["states"]
        lo, hi = np.asarray(P["o_space"]["low"], float), np.asarray(P["o_space"]["high"], float)
        a_lo, a_hi = np.asarray(P["a_space"]["low"], float), np.asarray(P["a_space"]["high"], float)
        u = np.asarray(u, float)
        if not hasattr(env, "u_prev"):
            env.u_prev = u
        cost = 0.0
        for k, sp in P["SP"].items():
            i = names.index(k)
            e = (x[i] - sp[env.t]) / (hi[i] - lo[i])
            cost += e ** 2 * P.get("r_scale", {}).get(k, 1)
        un, up = (u - a_lo) / (a_hi - a_lo), (env.u_prev - a_lo) / (a_hi - a_lo)
        cost += np.sum(R * ((un - up) if delta_u else un) ** 2)
        env.u_prev = u
        return -float(cost)
    return reward


def cstr_ocp():
    """pc-gym_paper/train_policies/cstr/cstr_train.py:14-49: CSTR with the OCP tracking custom reward."""
    p = _cstr_base()
    p.update(noise=True, noise_percentage=0.001, custom_reward=ocp_reward(0.1))
    return p


def first_order():
    """docs/env/first_order_system.md:38 action range; K=1, tau=0.5 defaults."""
    N = 30
    return {
        "N": N, "tsim": 3.0, "SP": {"x": [5.0] * N},
        "o_space": {"low": np.array([0.0, 0.0]), "high": np.array([10.0, 10.0])},
        "a_space": {"low": np.array([0.0]), "high": np.array([10.0])},
        "x0": np.array([0.25, 5.0]), "model": "first_order_system",
    }


def four_tank():
    """pc-gym_paper/train_policies/four_tank/4tank_train.py:54-85 (two setpoints), plus an affine
    constraint so the per-SP-key penalty (pcgym.py:552-557) is exercised."""
    N = 24
    SP = {"h3": [0.5] * (N // 2) + [0.1] * (N // 2), "h4": [0.2] * (N // 2) + [0.3] * (N // 2)}
    return {
        "N": N, "tsim": N * 1000 / 60, "SP": SP,
        "o_space": {"low": np.array([0.0] * 6), "high": np.array([0.6] * 6)},
        "a_space": {"low": np.array([0.1, 0.1]), "high": np.array([10.0, 10.0])},
        "x0": np.array([0.141, 0.112, 0.072, 0.42, SP["h3"][0], SP["h4"][0]]),
        "model": "four_tank", "noise": True, "noise_percentage": 0.05,
        "constraints": lambda x, u: np.array([x[0] - 0.2, 0.01 - x[2]]).reshape(-1,),
        "done_on_cons_vio": False, "r_penalty": True,
        "normalise_a": True, "normalise_o": False,
    }


def crystallization():
    """pc-gym_paper/train_policies/crystalisation/cryst_train.py:51-100: a_delta mode, two setpoints."""
    N = 30
    mu = [1478.00986666666, 22995.8230590611, 1800863.24079725, 248516167.940593]
    CV_0 = np.sqrt(mu[2] * mu[0] / (mu[1] ** 2) - 1)
    Ln_0 = mu[1] / (mu[0] + 1e-6)
    return {
        "N": N, "tsim": 30, "SP": {"CV": [1] * N, "Ln": [15] * N},
        "o_space": {"low": np.array([0, 0, 0, 0, 0, 0, 0, 0.9, 14.0]),
                    "high": np.array([1e20, 1e20, 1e20, 1e20, 0.5, 2, 20, 1.1, 16])},
        "a_space": {"low": np.array([-1.0]), "high": np.array([1.0])},
        "x0": np.array(mu + [0.15861523304, CV_0, Ln_0, 1, 15]),
        "model": "crystallization", "normalise_a": True, "normalise_o": True,
        "noise": True, "noise_percentage": 0.001,
        "a_0": 39, "a_delta": True, "a_space_act": {"low": np.array([10.0]), "high": np.array([40.0])},
    }


def four_tank_ocp():
    """pc-gym_paper/train_policies/four_tank/4tank_train.py:54-85 as written: two setpoints, 5 % noise and the
    script's own `oracle_reward` (:16-52, the OCP tracking reward with R = 0.1) as custom_reward.  BASELINE config 5's
    four_tank shard.  (N = 60, tsim = 1000 in the script; 24 intervals of the same dt here to keep the fixture small.)"""
    N = 24
    SP = {"h3": [0.5] * (N // 2) + [0.1] * (N // 2), "h4": [0.2] * (N // 2) + [0.3] * (N // 2)}
    return {
        "N": N, "tsim": N * 1000 / 60, "SP": SP,
        "o_space": {"low": np.array([0.0] * 6), "high": np.array([0.6] * 6)},
        "a_space": {"low": np.array([0.1, 0.1]), "high": np.array([10.0, 10.0])},
        "x0": np.array([0.141, 0.112, 0.072, 0.42, SP["h3"][0], SP["h4"][0]]),
        "model": "four_tank", "normalise_a": True, "normalise_o": True,
        "noise": True, "noise_percentage": 0.05,
        "custom_reward": ocp_reward(0.1),
    }


def cryst_reward(R=0.01):
    """The crystallization script's reward (pc-gym_paper/train_policies/crystalisation/cryst_train.py:17-48, written
    independently): CV and Ln are DERIVED from the noisy moments, normalised with o_space[5:7], tracked against
    SP['CV'][t], SP['Ln'][t]; plus R * (change of the a_space-normalised input)^2."""
    def reward(env, x, u, violated):
        P = env.env_params
        lo, hi = np.asarray(P["o_space"]["low"], float)[[5, 6]], np.asarray(P["o_space"]["high"], float)[[5, 6]]
        a_lo, a_hi = np.asarray(P["a_space"]["low"], float), np.asarray(P["a_space"]["high"], float)
        u = np.asarray(u, float)
        if not hasattr(env, "u_prev"):
            env.u_prev = u
        cv = (x[2] * x[0] / x[1] ** 2 - 1) ** 0.5
        ln = x[1] / x[0]
        e_cv = (P["SP"]["CV"][env.t] - cv) / (hi[0] - lo[0])
        e_ln = (P["SP"]["Ln"][env.t] - ln) / (hi[1] - lo[1])
        un, up = (u - a_lo) / (a_hi - a_lo), (env.u_prev - a_lo) / (a_hi - a_lo)
        env.u_prev = u
        return -float(e_cv ** 2 + e_ln ** 2 + np.sum(R * (un - up) ** 2))
    return reward


def cryst_ocp():
    """cryst_train.py:51-100 as written: a_delta mode, two setpoints AND the script's own reward.  BASELINE config 5's
    crystallization shard."""
    p = crystallization()
    p["custom_reward"] = cryst_reward(0.01)
    return p


def extraction():
    """pc-gym_paper/train_policies/multistage_extraction/me_train.py:54-93 with an X0 disturbance."""
    N = 20
    return {
        "N": N, "tsim": N, "SP": {"X5": [0.3] * (N // 4) + [0.4] * (N // 2) + [0.3] * (N - N // 4 - N // 2)},
        "o_space": {"low": np.array([0.0] * 10 + [0.3]), "high": np.array([1.0] * 10 + [0.4])},
        "a_space": {"low": np.array([5.0, 10.0]), "high": np.array([500.0, 1000.0])},
        "x0": np.array([0.55, 0.3, 0.45, 0.25, 0.4, 0.20, 0.35, 0.15, 0.25, 0.1, 0.3]),
        "model": "multistage_extraction", "r_scale": {"X5": 1},
        "disturbances": {"X0": np.random.default_rng(3).uniform(0.55, 0.65, N)},
        "disturbance_bounds": {"low": np.array([0.5]), "high": np.array([0.7])},
        "noise": True, "noise_percentage": 0.05,
    }


def biofilm():
    """pc-gym_paper/train_policies/Biofilm/biofilm_train.py:41-72 (no noise)."""
    N = 20
    return {
        "N": N, "tsim": 2 * N, "SP": {"S2_A": [1.5] * (N // 2) + [2.0] * (N // 2)},
        "o_space": {"low": np.array([0, 0, 0, 0] * 4 + [0.9]), "high": np.array([10, 10, 10, 500] * 4 + [1.1])},
        "a_space": {"low": np.array([0, 1, 0.05, 0.05, 0.05]), "high": np.array([10, 30, 1, 1, 1])},
        # documented x0 [2,0.1,10,0.1]*4 runs into the Monod pole S2 = -K2 within a few intervals
        # (SURVEY fact 9); start near the first state of the reference's own golden trajectory instead.
        "x0": np.array([11.0, 5.3, 3.7, 480.0] * 4 + [1.5]),
        "model": "biofilm_reactor", "noise": False,
    }


CASES = {
    "cstr_basic": cstr_basic,
    "cstr_dist_cons": cstr_dist_cons,
    "cstr_nonorm_cons": cstr_nonorm_cons,
    "cstr_one_dist": cstr_one_dist,
    "cstr_batch_reward": cstr_batch_reward,
    "cstr_ocp": cstr_ocp,
    "four_tank_ocp": four_tank_ocp,
    "cryst_ocp": cryst_ocp,
    "first_order": first_order,
    "four_tank": four_tank,
    "crystallization": crystallization,
    "extraction": extraction,
    "biofilm": biofilm,
}

ACTION_RANGE = {}


def actions_for(name, params, seed=0):
    """Deterministic action sequence (N-1 steps, reference episode length) in action_space units."""
    rng = np.random.default_rng(1000 + seed)
    n_a = np.asarray(params["a_space"]["low"]).shape[0]
    K = params["N"] - 1
    if params.get("normalise_a", True):
        lo, hi = ACTION_RANGE.get(name, (-1.0, 1.0))
        return rng.uniform(lo, hi, (K, n_a))
    lo, hi = np.asarray(params["a_space"]["low"], float), np.asarray(params["a_space"]["high"], float)
    return rng.uniform(lo, hi, (K, n_a))
