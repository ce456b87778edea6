"""CPU baseline: the restated reference step() timed on host cores.  TEST / BENCH INFRASTRUCTURE.

What is timed is `oracle.env.OracleEnv.step` (the reference's make_env.step restated line by line) with
scipy LSODA at rtol 1e-6 / atol 1e-8 standing in for CasADi's CVODES defaults (CasADi cannot be installed
in this image).  One worker process per host core, each stepping its own slice of the same
synthetic CSTR workload the GPU arm uses, one env at a time -- the reference has no batching.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def cstr_headline_params(N=100):
    """BASELINE config 3 for ONE env, reference-style (numpy) env_params.  SURVEY 8(d)."""
    return {
        "model": "cstr", "N": N, "tsim": 25.0,
        "SP": {"Ca": [0.85] * (N // 2) + [0.9] * (N - N // 2)},
        "o_space": {"low": np.array([0.7, 300.0, 0.8]), "high": np.array([1.0, 350.0, 0.9])},
        "a_space": {"low": np.array([295.0]), "high": np.array([302.0])},
        "x0": np.array([0.8, 330.0, 0.8]),
        "r_scale": {"Ca": 1e3},
        "noise": True, "noise_percentage": 0.001,
        "constraints": lambda x, u: np.array([319 - x[1], x[1] - 331]).reshape(-1,),
        "done_on_cons_vio": False, "r_penalty": True,
        "constraint_units": "physical",        # as the GPU arm: g sees physical T (bench.py cstr_headline)
        "normalise_a": True, "normalise_o": True,
    }


def _worker(args):
    wid, n_envs, n_steps, seed, integrator = args
    from oracle.env import OracleEnv
    rng = np.random.default_rng(seed + wid)
    params = cstr_headline_params()
    steps = 0
    t0 = time.perf_counter()
    for e in range(n_envs):
        p = dict(params)
        p["x0"] = np.array([0.8, 330.0, 0.8]) * np.concatenate([1 + 0.02 * rng.uniform(-1, 1, 2), [1.0]])
        env = OracleEnv(p, integrator=integrator)
        env.reset()
        acts = rng.uniform(-1, 1, (n_steps, 1))
        for k in range(n_steps):
            env.step(acts[k])
            steps += 1
    return steps, time.perf_counter() - t0


def run(envs_per_worker=8, n_steps=99, cores=None, seed=0, integrator="lsoda"):
    """Returns dict(value=env-steps/s aggregate, cores, sample, seconds)."""
    cores = cores or len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (cores or os.cpu_count())
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_worker, [(w, envs_per_worker, n_steps, seed, integrator) for w in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(s for s, _ in res)
    busy = max(t for _, t in res)
    return {
        "value": total / busy, "unit": "env-steps/s", "cores": cores, "kind": "port",
        "sample": f"{cores} workers x {envs_per_worker} CSTR envs x {n_steps} steps (config 3, one env per call, "
                  f"scipy LSODA rtol 1e-6/atol 1e-8 as CVODES-default stand-in); {total} env-steps in {busy:.1f}s busy",
        "per_core": total / busy / cores, "seconds": wall, "env_steps": total,
    }
