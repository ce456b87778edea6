#!/usr/bin/env python
"""Small rollouts through every kernel family (straight-line, warp-specialised pairs, per-SM incl. waves and the general
variant, generic adaptive, lane-split) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool synccheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgym_b200 import _native as nat  # noqa: E402
from pcgym_b200 import make_env  # noqa: E402

N = 14
base = {"model": "cstr", "N": N, "tsim": N * 0.25, "SP": {"Ca": [0.85] * N},
        "o_space": {"low": np.array([0.7, 300.0, 0.8]), "high": np.array([1.0, 350.0, 0.9])},
        "a_space": {"low": np.array([295.0]), "high": np.array([302.0])}, "x0": np.array([0.8, 330.0, 0.85]),
        "r_scale": {"Ca": 1e3}, "noise": True, "noise_percentage": 0.001,
        "constraints": lambda x, u: np.array([319 - x[1], x[1] - 331]).reshape(-1,), "done_on_cons_vio": False,
        "r_penalty": True, "constraint_units": "physical",
        "integrator": {"scheme": "tsit5", "adaptive": False, "n_sub": 1}}
rng = np.random.default_rng(0)
n_gen = 20000 + 6
dist_shared = {"disturbances": {"Ti": np.repeat(rng.uniform(340, 352, (N + 3) // 4), 4)[:N], "Caf": rng.uniform(0.9, 1.1, N)},
               "disturbance_bounds": {"low": np.array([330.0, 0.8]), "high": np.array([370.0, 1.2])},
               "o_space": {"low": np.array([0.7, 300.0, 0.8]), "high": np.array([1.0, 350.0, 0.9])},
               "integrator": {"scheme": "tsit5", "adaptive": False, "n_sub": 2, "check": False}}
dist_env = dict(dist_shared, disturbances={"Ti": rng.uniform(340, 352, (n_gen, N)), "Caf": rng.uniform(0.9, 1.1, (n_gen, N))})
ext = {"model": "multistage_extraction", "SP": {"Y1": [0.3] * N}, "r_scale": {"Y1": 3.0}, "tsim": float(N),
       "x0": np.array([0.55, 0.3, 0.45, 0.25, 0.4, 0.20, 0.35, 0.15, 0.25, 0.1, 0.3]),
       "a_space": {"low": np.array([5.0, 10.0]), "high": np.array([500.0, 1000.0])},
       "o_space": {"low": np.zeros(11), "high": np.ones(11)},
       "constraints": lambda x, u: np.array([0.5 - x[0], x[0] - 0.56]).reshape(-1,),
       "integrator": {"scheme": "tsit5", "adaptive": True, "rtol": 1e-6, "atol": 1e-9}}
base["integrator"]["check"] = False                        # no accuracy-probe launches under the sanitizer
for n, extra in ((4096 + 2, {}), (20000 + 6, {}), (78002, {}),                       # ws, per-SM, per-SM in two waves
                 (1001, {"integrator": {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}}),
                 (n_gen, dist_shared), (n_gen, dist_env),                            # general per-SM variant
                 (1003, ext), (7, dict(ext, model="multistage_extraction_10stage",   # lane-split kernel, partial warps / groups
                                      x0=np.concatenate([np.linspace(0.55, 0.1, 20), [0.3]]),
                                      o_space={"low": np.zeros(21), "high": np.ones(21)}))):
    env = make_env(dict(base, n_envs=n, **extra))
    env.reset()
    a = torch.rand(N - 1, env._a_lo.shape[0], n, dtype=torch.float64, device="cuda") * 2 - 1
    out = env.rollout(a)
    torch.cuda.synchronize()
    print(n, nat.lib().pcg_last_kernel().decode().split(" ")[0], float(out["reward"].sum()), int(env.status.max()))
