"""BASELINE config 5 (mixed CSTR / four_tank / crystallization population) through pcgym_b200.MixedBatch, against the
oracle env: the parts are bench.py's own definitions (bench.config5_parts) at a small batch size."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cases  # noqa: E402
import noise_model as nm  # noqa: E402
from oracle.env import OracleEnv  # noqa: E402

pytestmark = pytest.mark.gpu

SCRIPT_REWARD = {"four_tank": lambda: cases.ocp_reward(0.1), "crystallization": lambda: cases.cryst_reward(0.01)}


def _single_env_params(p, i):
    """The reference's env_params for env i of a batched part: host arrays, row i of every per-env table."""
    q = {k: v for k, v in p.items() if k not in ("n_envs", "device", "integrator", "seed", "env_id0", "constraint_units")}
    x0 = p["x0"]
    x0 = x0.detach().cpu().numpy() if isinstance(x0, torch.Tensor) else np.asarray(x0, float)
    q["x0"] = (x0[i] if x0.ndim == 2 else x0).copy()
    if p.get("disturbances"):
        q["disturbances"] = {k: (np.asarray(v)[i] if np.asarray(v).ndim == 2 else np.asarray(v)) for k, v in p["disturbances"].items()}
    if "constraint_units" in p:
        q["constraint_units"] = p["constraint_units"]
    return q


@pytest.mark.parametrize("part", ["cstr", "four_tank", "crystallization"])
def test_config5_parts_against_the_oracle_env(part):
    """Every part of config 5 EXACTLY as bench.py --config 5 defines it (integrator included), 256 envs per part in one
    MixedBatch.rollout: for a handful of envs the restated reference env (oracle/env.py, converged DOP853) is stepped
    through the same actions with the same Gaussian draws (the documented Philox draw replayed through its `normal`
    hook) and the scripts' own Python rewards: states to the north-star rtol 1e-6, observations, rewards, a_save."""
    import bench
    from pcgym_b200 import MixedBatch
    n = 256
    parts, acts = bench.config5_parts({"cstr": n, "four_tank": n, "crystallization": n}, device="cuda", seed=3)
    mb = MixedBatch(parts)
    mb.reset()
    out = mb.rollout(acts)
    torch.cuda.synchronize()
    assert mb.status_nonzero() == 0
    env, p, a = mb.envs[part], parts[part], acts[part].cpu().numpy()
    K, nx = a.shape[0], env.Nx_oracle
    got_x, got_obs, got_r = (out[part][k].cpu().numpy() for k in ("x", "obs", "reward"))
    key = env._desc.seed
    for i in (0, 1, n // 2 + 1, n - 1):
        q = _single_env_params(p, i)
        if part in SCRIPT_REWARD:
            q["custom_reward"] = SCRIPT_REWARD[part]()
        clock = {"t": 0}

        def normal(size, i=i, clock=clock):                  # called once per step with size = Nx (float percentage)
            clock["t"] += 1
            return nm.normals(key, np.array([i]), clock["t"], nx)[:size, 0]
        ref = OracleEnv(copy.deepcopy(q), integrator="truth", normal=normal)
        ref.reset()
        for k in range(K):
            obs, r, done, _, _ = ref.step(a[k, :, i].copy())
            np.testing.assert_allclose(got_x[k, :, i], ref.state[:nx], rtol=1e-6, atol=1e-12, err_msg=f"{part} env {i} x {k}")
            np.testing.assert_allclose(got_obs[k, :, i], obs, rtol=2e-6, atol=2e-6, err_msg=f"{part} env {i} obs {k}")
            np.testing.assert_allclose(got_r[k, i], r, rtol=1e-5, atol=1e-8, err_msg=f"{part} env {i} reward {k}")
        assert done and bool(out[part]["done"][-1, i])


def test_mixed_batch_is_the_parts_run_alone():
    """One MixedBatch.rollout (three streams) == the three make_env rollouts one after the other, bit for bit; a second
    pass from rewind(to_x0=True) with the same episode reproduces the first."""
    import bench
    from pcgym_b200 import MixedBatch, make_env
    sizes = {"cstr": 1024, "four_tank": 512, "crystallization": 258}
    parts, acts = bench.config5_parts(sizes, device="cuda", seed=1)
    mb = MixedBatch(parts)
    mb.reset(episode=0)
    out = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in mb.rollout(acts).items()}
    for k in mb.names:
        e = make_env(parts[k])
        e.reset(episode=0)
        alone = e.rollout(acts[k])
        for kk in ("x", "obs", "reward", "done", "violated"):
            assert torch.equal(out[k][kk], alone[kk]), (k, kk)
    mb.rewind(to_x0=True, episode=0)
    again = mb.rollout(acts)
    for k in mb.names:
        for kk in ("x", "obs", "done"):
            assert torch.equal(out[k][kk], again[k][kk]), (k, kk)
        # the OCP rewards carry u_prev across episodes like the scripts' `self.u_prev` does: only the first interval differs
        assert torch.equal(out[k]["reward"][1:], again[k]["reward"][1:]), k
    assert mb.status_nonzero() == 0


def test_mixed_batch_host_buffers():
    """rollout_host of every part (pinned host actions in; episode return, final observation and done flags back) ==
    the device rollout of the same episode."""
    import bench
    from pcgym_b200 import MixedBatch
    sizes = {"cstr": 512, "four_tank": 512, "crystallization": 256}
    parts, acts = bench.config5_parts(sizes, device="cuda", seed=2)
    mb = MixedBatch(parts)
    mb.reset()
    mb.rollout(acts)                                          # from here on u_prev of the OCP rewards is valid
    mb.rewind(to_x0=True, episode=5)
    dev = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in mb.rollout(acts).items()}
    mb.rewind(to_x0=True, episode=5)
    host = mb.rollout_host({k: a.cpu().pin_memory() for k, a in acts.items()}, outputs=("return", "obs_last", "done"))
    for k in mb.names:
        np.testing.assert_allclose(host[k]["return"].numpy(), dev[k]["reward"].sum(0).cpu().numpy(), rtol=1e-12, atol=1e-12)
        assert torch.equal(host[k]["obs_last"], dev[k]["obs"][-1].cpu())
        assert torch.equal(host[k]["done"], dev[k]["done"].cpu())
    assert mb.status_nonzero() == 0
