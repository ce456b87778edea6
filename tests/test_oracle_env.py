"""oracle.env.OracleEnv vs per-step outputs of the reference's own make_env (tests/golden/step_*.npz)."""
import os

import numpy as np
import pytest

import cases
from oracle.env import OracleEnv


def replay(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"step_{name}.npz"))
    params = cases.CASES[name]()
    zq = []
    env = OracleEnv(params, integrator="truth", normal=lambda size: zq.pop(0))
    obs0, info = env.reset()
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-12, atol=1e-12)
    names = env.model.info()["states"]
    for k in range(g["actions"].shape[0]):
        z = g["z"][k]
        if not np.all(np.isnan(z)):
            if isinstance(params.get("noise_percentage"), float):
                zq.append(z.copy())
            else:
                zq.extend(np.array([z[i]]) for i in range(env.Nx_oracle) if names[i] in params["noise_percentage"])
        obs, r, done, trunc, info = env.step(g["actions"][k].copy())
        np.testing.assert_allclose(env.state, g["state"][k], rtol=1e-9, atol=1e-12, err_msg=f"state step {k}")
        np.testing.assert_allclose(obs, g["obs"][k], rtol=1e-8, atol=1e-9, err_msg=f"obs step {k}")
        np.testing.assert_allclose(info["obs"], g["info_obs"][k], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(r, g["reward"][k], rtol=1e-7, atol=1e-10, err_msg=f"reward step {k}")
        assert bool(done) == bool(g["done"][k]) and trunc is False
        if env.a_delta:
            np.testing.assert_allclose(np.atleast_1d(env.a_save), g["a_save"][k], rtol=1e-12)
    if "cons_info" in g.files:
        np.testing.assert_allclose(info["cons_info"], g["cons_info"], rtol=1e-8, atol=1e-9)
    assert not zq


@pytest.mark.parametrize("name", list(cases.CASES))
def test_step_semantics_match_reference(name, golden_dir):
    replay(name, golden_dir)


def test_episode_is_N_minus_1_steps():
    env = OracleEnv(cases.first_order(), integrator="dense")
    env.reset()
    n = 0
    done = False
    while not done:
        _, _, done, _, _ = env.step(np.array([0.1]))
        n += 1
    assert n == env.N - 1


def test_bad_params():
    with pytest.raises(ValueError):
        OracleEnv([1, 2])
    p = cases.first_order()
    p["model"] = "nope"
    with pytest.raises(ValueError):
        OracleEnv(p)
