"""Row f1: batched `policy_eval.get_rollouts` (reference src/pcgym/policy_evaluation.py:71-197) on the CUDA path
against the reference's rollout loop restated over the oracle env; device policies, CUDA-graph replay,
`make_env.get_rollouts`, and `reproducibility_metric` on the resulting dictionary."""
import numpy as np
import pytest
import torch

import cases
from oracle.env import OracleEnv

pytestmark = pytest.mark.gpu

TIGHT = {"scheme": "dopri5", "adaptive": False, "n_sub": 48}


class HostP:
    """Stable-Baselines3 style policy: numpy in, (action, state) out; P-controller on the normalised observation."""
    calls = 0

    def predict(self, obs, deterministic=True):
        HostP.calls += 1
        o = np.asarray(obs, dtype=np.float64)
        a = np.clip(5.0 * (o[..., 2:3] - o[..., 0:1]), -1, 1)
        return a, None


def dev_policy(obs):
    return torch.clamp(5.0 * (obs[:, 2:3] - obs[:, 0:1]), -1, 1)


def reference_rollout(params, policy):
    """policy_evaluation.py:71-130 verbatim in structure, over the oracle env."""
    env = OracleEnv(params)
    lo, hi = env.observation_space_base.low, env.observation_space_base.high
    alo, ahi = params["a_space"]["low"], params["a_space"]["high"]
    total_reward = []
    s = np.zeros((env.Nx, env.N))
    acts = np.zeros((alo.shape[0], env.N))
    partial = bool(params.get("partial_observation"))
    o, info = env.reset()
    total_reward.append(info["r_init"])
    s[:, 0] = ((info["obs"] if partial else o) + 1) * (hi - lo) / 2 + lo          # policy_evaluation.py:90-98
    for i in range(env.N - 1):
        a, _ = policy.predict(o, deterministic=True)
        o, r, term, trunc, info = env.step(a)
        acts[:, i] = (a + 1) * (ahi - alo) / 2 + alo
        s[:, i + 1] = ((info["obs"] if partial else o) + 1) * (hi - lo) / 2 + lo   # policy_evaluation.py:107-114
        total_reward.append(r)
    cons = info["cons_info"] if env.constraint_active else np.zeros((1, env.N, 1))
    a, _ = policy.predict(o, deterministic=True)
    acts[:, env.N - 1] = (a + 1) * (ahi - alo) / 2 + alo
    return np.array(total_reward, dtype=float), s, acts, cons


def _params():
    p = cases.cstr_dist_cons()
    p.update(noise=False, integrator=TIGHT)
    return p


def test_get_rollouts_matches_reference_loop_shapes_and_values():
    from pcgym_b200 import make_env, policy_eval
    p, reps = _params(), 5
    r_ref, s_ref, a_ref, g_ref = reference_rollout(p, HostP())
    ev = policy_eval(make_env, {"host": HostP(), "dev": dev_policy}, reps, p)
    assert ev.n_pi == 2 and ev.reps == reps and not ev.oracle and ev.env.n_envs == reps
    data = ev.get_rollouts()
    N, Nx = p["N"], 5
    for name in ("host", "dev"):
        d = data[name]
        assert d["r"].shape == (1, N, reps) and d["x"].shape == (Nx, N, reps)
        assert d["u"].shape == (1, N, reps) and d["g"].shape == (2, N, 1, reps)
        for k in range(reps):                        # noise off: every repetition is the reference episode
            np.testing.assert_allclose(d["x"][:, :, k], s_ref, rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(d["u"][:, :, k], a_ref, rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(d["r"][0, :, k], r_ref, rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(d["g"][:, :, :, k], g_ref, rtol=1e-6, atol=1e-7)
    assert ev.data is data


def test_get_rollouts_with_partial_observation_records_the_full_state():
    """partial_observation hides states from the POLICY (the returned observation has them zeroed, pcgym.py:495-498),
    but the rollout collector records the full state from info["obs"] (policy_evaluation.py:94-98, 110-114): the hidden
    state's trajectory in data["x"] is the real one, not the observation-space midpoint (round-1 ADVICE)."""
    from pcgym_b200 import make_env, policy_eval

    class TP:                                                  # acts on the temperature: Ca is hidden from it
        def predict(self, obs, deterministic=True):
            o = np.asarray(obs, dtype=np.float64)
            return np.clip(-3.0 * o[..., 1:2], -1, 1), None

    p = cases.cstr_one_dist()
    assert p["partial_observation"] == ["Ca"]
    p.update(noise=False, integrator=TIGHT)
    r_ref, s_ref, a_ref, _ = reference_rollout(p, TP())
    assert np.ptp(s_ref[0]) > 1e-3                             # the hidden state moves
    reps = 3
    data = policy_eval(make_env, {"host": TP(), "dev": lambda obs: torch.clamp(-3.0 * obs[:, 1:2], -1, 1)}, reps, p).get_rollouts()
    for name in ("host", "dev"):
        for k in range(reps):
            np.testing.assert_allclose(data[name]["x"][:, :, k], s_ref, rtol=1e-6, atol=1e-9, err_msg=name)
            np.testing.assert_allclose(data[name]["u"][:, :, k], a_ref, rtol=1e-6, atol=1e-9, err_msg=name)
            np.testing.assert_allclose(data[name]["r"][0, :, k], r_ref, rtol=1e-5, atol=1e-9, err_msg=name)


def test_cuda_graph_replay_equals_eager_and_repeats():
    from pcgym_b200 import make_env, policy_eval
    p = cases.cstr_basic()                           # Philox noise on: repetitions differ, runs are reproducible
    p.update(integrator={"scheme": "tsit5", "adaptive": False, "n_sub": 1}, seed=3)
    reps = 256
    eager = policy_eval(make_env, {"dev": dev_policy}, reps, p, as_torch=True).get_rollouts()["dev"]
    ev = policy_eval(make_env, {"dev": dev_policy}, reps, p, cuda_graph=True, as_torch=True)
    first = {k: v.clone() for k, v in ev.get_rollouts()["dev"].items()}
    second = ev.get_rollouts()["dev"]                # replay of the captured episode
    for k in ("r", "x", "u"):
        assert torch.equal(first[k], eager[k]), k
        assert torch.equal(second[k], eager[k]), k
    assert float(eager["x"][0].std(dim=-1).max()) > 0          # noisy repetitions are not identical


def test_env_get_rollouts_single_rep_host_policy_and_metric():
    from pcgym_b200 import make_env, reproducibility_metric
    p = _params()
    env = make_env(p)
    HostP.calls = 0
    evaluator, data = env.get_rollouts({"pi": HostP()}, reps=1)
    assert HostP.calls == p["N"]                     # N-1 steps + the trailing predict
    r_ref, s_ref, a_ref, _ = reference_rollout(p, HostP())
    np.testing.assert_allclose(data["pi"]["x"][:, :, 0], s_ref, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(data["pi"]["r"][0, :, 0], r_ref, rtol=1e-5, atol=1e-9)
    m = reproducibility_metric("std", "mean", -1.0).evaluate(evaluator)
    assert m["pi"]["r"].shape == (1, p["N"]) and m["pi"]["g"].shape == (p["N"], 1)
    with pytest.raises(NotImplementedError):
        env.get_rollouts({"pi": HostP()}, reps=1, oracle=True)
    with pytest.raises(NotImplementedError):
        env.plot_rollout({"pi": HostP()}, reps=1)
