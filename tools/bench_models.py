#!/usr/bin/env python
"""Rollout throughput of every registry model (BASELINE configs 2-4 + the rest), information only.

    python tools/bench_models.py [--steps 10]

Prints one JSON line per (model, integrator): env-steps/s, achieved GB/s on the algorithmic rollout bytes of
SURVEY 8(d) (shared setpoint schedule: 8 B less than the survey's per-env figure), fraction of the measured
HBM peak.  bench.py remains the contract benchmark (CSTR headline); BASELINE config 5 (the mixed population) is
`bench.py --config 5`.
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
from pcgym_b200 import _native as nat  # noqa: E402
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
    # the stiff integrator (Rosenbrock RODAS4, analytic Jacobian) on the same two workloads, same tolerances
    ("multistage_extraction", 32768, 59, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-10}),
    ("cstr", 65536, 99, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-8}),
    ("multistage_extraction", 32768, 59, {"scheme": "rodas4", "rtol": 1e-6, "atol": 1e-9}),
    ("multistage_extraction", 32768, 59, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-6, "atol": 1e-9}),
    # BASELINE config 4 as worded ("10-stage, ~20-state"): the labelled extension model
    ("multistage_extraction_10stage", 32768, 59, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}),
    ("multistage_extraction_10stage", 32768, 59, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-10}),
]


def bytes_per_env_step(env):
    nx, nu = env.Nx_oracle, env._a_lo.shape[0]
    return 8 * nu + 8 * (nx + env._n_obs + 1 + env.n_con) + 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=0, help="BASELINE config 2 (first_order) or 4 (extraction): only those rows")
    ap.add_argument("--only", default="", help="comma-separated indices into CONFIGS (ncu captures of one kernel family)")
    args = ap.parse_args()
    only = {int(v) for v in args.only.split(",") if v != ""} or {2: {0, 1}, 4: {10, 11, 15, 16}}.get(args.config, set())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for idx, cfg in enumerate(CONFIGS):
        if only and idx not in only:
            continue
        model, n, K, integ = cfg[:4]
        noise = cfg[4] if len(cfg) > 4 else True
        params, x0, acts, key, sp = T.batch_setup(model, n, K)
        params = dict(params, integrator=integ, noise=noise, noise_percentage=0.001)
        if model == "four_tank":              # pumps near their minimum for several intervals drain a tank (sqrt of a negative level)
            acts = acts * 0.75 + 0.25
        env = make_env(params)
        env.reset()
        a = torch.as_tensor(acts, device="cuda")
        # every pass starts from the reset() state (one device copy of x0): random inputs drive some models out of their
        # physical range within a few dozen intervals, and a pass over diverged lanes would time nothing meaningful
        for _ in range(args.warmup):
            env.rewind(to_x0=True); env.rollout(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            env.rewind(to_x0=True); env.rollout(a)
        e1.record(); torch.cuda.synchronize()
        bad = int((env.status != 0).sum())
        if bad:
            raise SystemExit(f"{model}: {bad} envs with a non-zero integrator status -- not a valid measurement")
        ms = e0.elapsed_time(e1) / args.steps
        b = bytes_per_env_step(env)
        rate = n * K / (ms * 1e-3)
        rhs = float(env.rhs_evaluations.double().mean()) / (K * (args.steps + args.warmup))
        print(json.dumps({"model": model, "n_envs": n, "intervals": K, "integrator": integ, "noise": noise,
                          "ms_per_launch": round(ms, 4),
                          "env_steps_per_s": rate, "bytes_per_env_step": b, "GBps": rate * b / 1e9,
                          "hbm_frac": rate * b / 1e9 / peak, "rhs_evals_per_interval": round(rhs, 1),
                          "kernel": nat.lib().pcg_last_kernel().decode(), "status_nonzero": bad}), flush=True)
        del env, a
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
