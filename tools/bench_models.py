#!/usr/bin/env python
"""Rollout throughput of every registry model (BASELINE configs 2-4 + the rest), information only.

    python tools/bench_models.py [--steps 10]

Prints one JSON line per (model, integrator): env-steps/s, achieved GB/s on the algorithmic rollout bytes of
SURVEY 8(d) (shared setpoint schedule: 8 B less than the survey's per-env figure), fraction of the measured
HBM peak.  bench.py remains the contract benchmark (CSTR headline).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pcgym_b200 import make_env  # noqa: E402
import test_gpu_parity as T  # noqa: E402  (reuses the per-model batch scenarios)

CONFIGS = [   # model, n_envs, K, integrator[, measurement noise (default on)]
    # BASELINE config 2 exactly (noise off, no constraints): the bandwidth sanity case
    ("first_order_system", 65536, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}, False),
    ("first_order_system", 1048576, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}, False),
    ("first_order_system", 65536, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}),
    ("first_order_system", 1048576, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}),
    ("cstr", 65536, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}),
    ("cstr", 524288, 99, {"scheme": "tsit5", "adaptive": False, "n_sub": 1}),
    ("cstr", 65536, 99, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}),
    ("four_tank", 65536, 59, {"scheme": "tsit5", "adaptive": False, "n_sub": 2}),
    ("crystallization", 65536, 19, {"scheme": "tsit5", "adaptive": False, "n_sub": 8}),   # random inputs diverge the moments after ~24 intervals (oracle too)
    ("biofilm_reactor", 32768, 19, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}),
    ("multistage_extraction", 32768, 59, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}),
]


def bytes_per_env_step(env):
    nx, nu = env.Nx_oracle, env._a_lo.shape[0]
    return 8 * nu + 8 * (nx + env._n_obs + 1 + env.n_con) + 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for cfg in CONFIGS:
        model, n, K, integ = cfg[:4]
        noise = cfg[4] if len(cfg) > 4 else True
        params, x0, acts, key, sp = T.batch_setup(model, n, K)
        params = dict(params, integrator=integ, noise=noise, noise_percentage=0.001)
        env = make_env(params)
        env.reset()
        a = torch.as_tensor(acts, device="cuda")
        for _ in range(3):
            env.rewind(); env.rollout(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            env.rewind(); env.rollout(a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        b = bytes_per_env_step(env)
        rate = n * K / (ms * 1e-3)
        rhs = float(env.rhs_evaluations.double().mean()) / (K * (args.steps + 3))
        print(json.dumps({"model": model, "n_envs": n, "intervals": K, "integrator": integ, "noise": noise,
                          "ms_per_launch": round(ms, 4),
                          "env_steps_per_s": rate, "bytes_per_env_step": b, "GBps": rate * b / 1e9,
                          "hbm_frac": rate * b / 1e9 / peak, "rhs_evals_per_interval": round(rhs, 1),
                          "status_nonzero": int((env.status != 0).sum())}), flush=True)
        del env, a
        torch.cuda.empty_cache()
    mixed(args, peak)


def mixed(args, peak):
    """BASELINE config 5 on ONE GPU's share: a third each cstr (live Ti / Caf disturbances per env per step),
    four_tank, crystallization (524 288 envs over 8 GPUs = 65 536 per GPU, every GPU carrying the same
    mix); the three rollouts are launched on three streams and timed together."""
    n_each, integ = 21846, {"scheme": "tsit5", "adaptive": False}
    envs, acts, Ks, byts = [], [], [], []
    for model, K, n_sub in (("cstr", 99, 1), ("four_tank", 59, 2), ("crystallization", 19, 8)):
        params, x0, a, key, sp = T.batch_setup(model, n_each, K)
        params = dict(params, integrator=dict(integ, n_sub=n_sub), noise=True, noise_percentage=0.001)
        if model == "cstr":
            rng = np.random.default_rng(1)
            params["disturbances"] = {"Ti": np.repeat(rng.uniform(340, 360, (n_each, (K + 1 + 3) // 4)), 4, axis=1)[:, :K + 1],
                                      "Caf": rng.uniform(0.9, 1.1, (n_each, K + 1))}
            params["disturbance_bounds"] = {"low": np.array([330.0, 0.8]), "high": np.array([370.0, 1.2])}
        env = make_env(params)
        env.reset()
        envs.append(env); acts.append(torch.as_tensor(a, device="cuda")); Ks.append(K)
        byts.append(bytes_per_env_step(env) + (16 if model == "cstr" else 0))      # + per-env disturbance reads
    streams = [torch.cuda.Stream() for _ in envs]

    def once():
        for env, a, s in zip(envs, acts, streams):
            with torch.cuda.stream(s):
                env.rewind(); env.rollout(a)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(args.steps):
        once()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    steps = sum(n_each * K for K in Ks)
    gb = sum(n_each * K * b for K, b in zip(Ks, byts)) / 1e9
    print(json.dumps({"model": "mixed cstr(+dist)/four_tank/crystallization (config 5, one GPU share)",
                      "n_envs": 3 * n_each, "intervals": Ks, "ms_per_pass": round(ms, 4),
                      "env_steps_per_s": steps / (ms * 1e-3), "GBps": gb / (ms * 1e-3),
                      "hbm_frac": gb / (ms * 1e-3) / peak}), flush=True)


if __name__ == "__main__":
    main()
