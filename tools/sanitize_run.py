#!/usr/bin/env python
"""Small rollouts through the three straight-line kernels and the generic one, for compute-sanitizer:

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
for n, extra in ((4096 + 2, {}), (20000 + 6, {}), (78002, {}),
                 (1001, {"integrator": {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}})):
    env = make_env(dict(base, n_envs=n, **extra))
    env.reset()
    a = torch.rand(N - 1, 1, n, dtype=torch.float64, device="cuda") * 2 - 1
    out = env.rollout(a)
    torch.cuda.synchronize()
    print(n, nat.lib().pcg_last_kernel().decode().split(" ")[0], float(out["reward"].sum()), int(env.status.max()))
