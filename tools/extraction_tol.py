"""Worst relative trajectory error of multistage_extraction under the adaptive Tsit5 pair against the converged solution,
per tolerance setting (which (rtol, atol) meets the north-star 1e-6 bar for BASELINE config 4, and what it costs)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from pcgym_b200 import make_env, _native as nat

model, n, K = "multistage_extraction", 384, 59
params, x0, acts, key, sp = T.batch_setup(model, n, K)
xs, _ = T.oracle_batch(model, params, x0, acts, key, sp, 4096)
for rtol, atol in ((1e-8, 1e-10), (1e-7, 1e-9), (3e-7, 1e-9), (1e-6, 1e-9), (1e-6, 1e-8)):
    env = make_env(dict(params, integrator={"scheme": "tsit5", "adaptive": True, "rtol": rtol, "atol": atol}))
    env.reset()
    out = env.rollout(torch.as_tensor(acts).cuda())
    err = T.rel(out["x"].cpu().numpy(), xs)
    print(f"rtol {rtol:g} atol {atol:g}: max rel err {err:.2e}, rhs/interval {float(env.rhs_evaluations.double().mean()) / K:.1f}, "
          f"{nat.lib().pcg_last_kernel().decode()[:24]}", flush=True)
