"""A batched affine custom model (the reference's custom_model hook, pcgym.py:150-153) through the Affine<3,2> kernels:
65 536 envs x 49 intervals, fixed Tsit5 -- the workload of the ncu capture under profiles/ (profiling aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgym_b200 import make_env, _native as nat


class Two:
    int_method = "casadi"

    def __call__(self, x, u):
        return np.array([-x[0] + 0.5 * x[2] + u[0] + 0.1, -2 * x[1] + u[1], x[0] - x[2] - 0.3])

    def info(self):
        return {"parameters": {}, "states": ["a", "b", "c"], "inputs": ["p", "q"], "disturbances": []}


n, K = 65536, 49
rng = np.random.default_rng(2)
x0 = rng.uniform(-1, 1, (n, 3))
p = {"custom_model": Two(), "n_envs": n, "N": K + 1, "tsim": (K + 1) * 0.2, "SP": {"b": [0.0] * (K + 1)},
     "x0": np.concatenate([x0, np.zeros((n, 1))], axis=1), "normalise_a": False, "normalise_o": False,
     "a_space": {"low": -np.ones(2), "high": np.ones(2)}, "o_space": {"low": -np.ones(4), "high": np.ones(4)},
     "noise": True, "noise_percentage": 0.001, "integrator": {"scheme": "tsit5", "adaptive": False, "n_sub": 1, "check": False}}
env = make_env(p)
env.reset()
a = torch.as_tensor(rng.uniform(-1, 1, (K, 2, n))).cuda()
for _ in range(4):
    env.rewind(to_x0=True)
    env.rollout(a)
torch.cuda.synchronize()
print(nat.lib().pcg_last_kernel().decode(), int(env.status.max()))
