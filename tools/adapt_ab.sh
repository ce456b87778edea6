#!/bin/bash
# same-box A/B of two library builds (pcgym_b200/lib_a.so, lib_b.so) on the ADAPTIVE path: CSTR and
# multistage_extraction rollouts with adaptive Tsit5; prints ms per launch and RHS evaluations per attempted step / 7
# (profiling aid)
for L in pcgym_b200/lib_a.so pcgym_b200/lib_b.so pcgym_b200/lib_a.so pcgym_b200/lib_b.so; do PCG_LIB=$PWD/$L python - <<'PY'
import sys, os, json, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from pcgym_b200 import make_env
import test_gpu_parity as T
for model, n, K, integ in [("cstr", 65536, 99, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}), ("multistage_extraction", 32768, 20, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10})]:
    params, x0, acts, key, sp = T.batch_setup(model, n, K)
    env = make_env(dict(params, integrator=integ, noise=True, noise_percentage=0.001)); env.reset()
    a = torch.as_tensor(acts, device="cuda")
    for _ in range(2): env.rewind(); env.rollout(a)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env.rewind(); env.rollout(a)
    e1.record(); torch.cuda.synchronize()
    print(os.environ["PCG_LIB"][-8:], model, round(e0.elapsed_time(e1) / 5, 4), "ms", float(env.rhs_evaluations.double().mean()) / (K * 7))
PY
done
