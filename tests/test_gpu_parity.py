"""CUDA path vs oracle / golden fixtures.  Everything here calls the kernels through the C ABI
(pcgym_b200.make_env -> ctypes -> libpcgym_b200.so).  Tolerance: north-star rtol 1e-6 on fp64
trajectories; tighter where both sides are converged."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import integrate as oint
from oracle import rhs as orhs

pytestmark = pytest.mark.gpu

TIGHT = {"scheme": "dopri5", "adaptive": False, "n_sub": 48}       # converged fixed-step setting (extraction is stiff)
HEADLINE = {"scheme": "tsit5", "adaptive": False, "n_sub": 1}      # what bench.py times


def _env(params, **over):
    from pcgym_b200 import make_env
    p = dict(params)
    p.update(over)
    return make_env(p)


def same(a, b):
    """bitwise equality that treats NaN == NaN (diverged envs must diverge identically)."""
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


def rel(a, b, floor=1e-12):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


# ----------------------------------------------------------------------------- step semantics vs reference fixtures
@pytest.mark.parametrize("name", list(cases.CASES))
def test_step_cases_match_reference_fixture(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"step_{name}.npz"))
    params = cases.CASES[name]()
    n_sub = 400 if name == "extraction" else 48
    env = _env(params, noise_mode="external", integrator=dict(TIGHT, n_sub=n_sub))
    obs0, info = env.reset()
    assert isinstance(obs0, np.ndarray) and obs0.shape == g["obs0"].shape
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-12, atol=1e-12)
    noisy = bool(params.get("noise", False))
    K = g["actions"].shape[0]
    for k in range(K):
        z = np.nan_to_num(g["z"][k], nan=0.0)[None] if noisy else None
        obs, r, done, trunc, info = env.step(g["actions"][k].copy(), z=z)
        assert isinstance(obs, np.ndarray) and isinstance(r, float) and isinstance(done, bool) and trunc is False
        np.testing.assert_allclose(env.state, g["state"][k], rtol=2e-8, atol=1e-11, err_msg=f"{name} state {k}")
        np.testing.assert_allclose(obs, g["obs"][k], rtol=5e-7, atol=5e-8, err_msg=f"{name} obs {k}")
        # info["obs"] is the normalised observation BEFORE the partial-observation mask (pcgym.py:489 vs :495-498)
        io_ = info["obs"]
        io_ = io_.detach().cpu().numpy() if isinstance(io_, torch.Tensor) else np.asarray(io_)
        np.testing.assert_allclose(io_.reshape(-1), g["info_obs"][k], rtol=5e-7, atol=5e-8, err_msg=f"{name} info obs {k}")
        np.testing.assert_allclose(r, g["reward"][k], rtol=1e-6, atol=1e-9, err_msg=f"{name} reward {k}")
        assert done == bool(g["done"][k]), f"{name} done {k}"
        if env.a_delta:
            np.testing.assert_allclose(env.a_save, g["a_save"][k], rtol=1e-12)
    if "cons_info" in g.files:
        ci = info["cons_info"].cpu().numpy()
        np.testing.assert_allclose(ci, g["cons_info"], rtol=1e-6, atol=1e-7)
    assert int(env.status.max()) == 0
    with pytest.raises(IndexError):
        env.step(g["actions"][0])


@pytest.mark.parametrize("name", ["cstr_dist_cons", "four_tank", "crystallization", "extraction"])
def test_rollout_equals_repeated_steps_bitwise(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"step_{name}.npz"))
    params = cases.CASES[name]()
    n = 64
    rng = np.random.default_rng(3)
    acts = torch.as_tensor(rng.uniform(-1, 1, (g["actions"].shape[0], g["actions"].shape[1], n)))
    e1 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e2 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e1.reset(); e2.reset()
    out = e1.rollout(acts)
    for k in range(acts.shape[0]):
        obs, r, done, trunc, info = e2.step(acts[k].T.cuda())
        assert same(obs.T, out["obs"][k]), f"obs step {k}"
        assert same(r, out["reward"][k])
        assert torch.equal(done, out["done"][k].bool())
    assert same(e1._x, e2._x)
    if e1.n_con:
        assert same(e1.info["cons_info"], e2.info["cons_info"])


@pytest.mark.parametrize("chunks", [1, 4, 7])
def test_rollout_host_pipelined_equals_rollout_bitwise(chunks):
    """make_env.rollout_host: time-sliced H2D / kernel / D2H pipeline over host buffers == one device rollout."""
    params = cases.cstr_dist_cons()
    n, K = 512, params["N"] - 1
    rng = np.random.default_rng(4)
    acts = torch.as_tensor(rng.uniform(-1, 1, (K, 1, n)))
    e1 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e2 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e1.reset(); e2.reset()
    ref = e1.rollout(acts.cuda())
    out = e2.rollout_host(acts.pin_memory(), chunks=chunks, want_x=True)
    for k in ("obs", "reward", "done", "x"):
        assert out[k].device.type == "cpu" and out[k].is_pinned()
        assert same(out[k].to(torch.float64), ref[k].cpu().to(torch.float64)), k
    assert e2.t == K and torch.equal(e1.info["cons_info"], e2.info["cons_info"])
    e2.reset(episode=0)
    out2 = e2.rollout_host(acts, out=out, chunks=chunks, want_x=True)
    assert out2 is out and same(out["obs"], ref["obs"].cpu())
    with pytest.raises(ValueError):
        e2.rollout_host(acts.cuda())
    # back-to-back calls without waiting: two result sets alternate, host_sync() completes both
    a2 = torch.as_tensor(rng.uniform(-1, 1, (K, 1, n))).pin_memory()
    e1.reset(episode=3); ref1 = {k: v.clone() for k, v in e1.rollout(acts.cuda()).items() if v is not None}
    e1.rewind(); ref2 = e1.rollout(a2.cuda())
    e2.reset(episode=3)
    o1 = e2.rollout_host(acts.pin_memory(), chunks=chunks, wait=False)
    e2.rewind()
    o2 = e2.rollout_host(a2, chunks=chunks, wait=False)
    e2.host_sync()
    assert o1 is not o2 and same(o1["obs"], ref1["obs"].cpu()) and same(o2["obs"], ref2["obs"].cpu())
    assert same(o2["reward"], ref2["reward"].cpu())


def test_rollout_host_selected_outputs():
    """rollout_host(outputs=...): reduced result sets (episode return reduced on the device, final observation,
    flags only) equal the corresponding slices / sums of the full device rollout."""
    params = cases.cstr_dist_cons()
    n, K = 512, params["N"] - 1
    acts = torch.as_tensor(np.random.default_rng(6).uniform(-1, 1, (K, 1, n)))
    e1 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e2 = _env(params, n_envs=n, integrator=HEADLINE, seed=5)
    e1.reset(); e2.reset()
    ref = e1.rollout(acts.cuda(), copy=True)
    out = e2.rollout_host(acts.pin_memory(), chunks=5, outputs=("return", "obs_last", "done", "violated"))
    assert set(out) == {"return", "obs_last", "done", "violated"} and all(v.is_pinned() for v in out.values())
    np.testing.assert_allclose(out["return"].numpy(), ref["reward"].sum(0).cpu().numpy(), rtol=1e-13)
    assert same(out["obs_last"], ref["obs"][-1].cpu())
    assert torch.equal(out["done"], ref["done"].cpu()) and torch.equal(out["violated"], ref["violated"].cpu())
    with pytest.raises(ValueError):
        e2.rollout_host(acts, outputs=("obs", "nope"))
    # rollout(copy=True) owns its result: a second rollout does not overwrite it
    e1.rewind(episode=0)
    keep = ref["obs"].clone()
    e1.rollout(torch.zeros_like(acts).cuda())
    assert same(ref["obs"], keep)


@pytest.mark.parametrize("n,request_", [(4096, "auto"), (20000, "auto"), (78000, "fast"), (78000, "auto")])
def test_split_rollouts_equal_one_rollout_bitwise(n, request_):     # pair-per-CTA, per-SM pool (partial CTA), monolithic, per-SM pool in two waves
    """A K-interval rollout cut at arbitrary points (1, 2, 5, 8, 9 intervals: below / at / above the ring depth of the
    warp-specialised kernels, t0 > 0) reproduces the single launch bit for bit, including the sticky done flag and
    cons_info rows."""
    params = cases.cstr_dist_cons()
    params.pop("disturbances"); params.pop("disturbance_bounds")
    params.update(noise=True, noise_percentage=0.001, integrator=HEADLINE, seed=3)
    K = params["N"] - 1
    rng = np.random.default_rng(12)
    acts = torch.as_tensor(rng.uniform(-1, 1, (K, 1, n))).cuda()
    one, many = _env(params, n_envs=n), _env(params, n_envs=n)
    one.reset(); many.reset()
    with forced_kernel(request_):
        ref = one.rollout(acts)
        cuts, k0 = [1, 2, 5, 8, 9], 0
        parts = []
        for c in cuts + [K - sum(cuts)]:
            out = many.rollout(acts[k0:k0 + c])
            parts.append({k: v.clone() for k, v in out.items() if v is not None})
            k0 += c
    for key in ("x", "obs", "reward", "done", "violated"):
        assert same(torch.cat([p[key] for p in parts]).double(), ref[key].double()), key
    assert same(one.info["cons_info"], many.info["cons_info"]) and same(one._x, many._x)
    assert torch.equal(one._done_state, many._done_state) and bool(one._done_state.all())


def test_three_straight_line_kernels_agree():
    """The batch size picks the kernel (csrc/launch.cuh): warp-specialised pairs (n = 4 096), persistent CTA per SM with
    a finisher pool (n = 20 000, last CTA partial), software-pipelined 255-register kernel (n = 78 000).  The first
    4 096 envs get the same x0 / actions / noise stream in all three, so their outputs must agree to rounding (the
    integrator warps use a Horner-form exp, the monolithic kernel an Estrin-form one: not bitwise)."""
    from pcgym_b200 import _native as nat
    params = cases.cstr_dist_cons()
    params.pop("disturbances"); params.pop("disturbance_bounds")
    params.update(noise=True, noise_percentage=0.001, integrator=HEADLINE, seed=21, x0=None)
    K, m = params["N"] - 1, 4096
    rng = np.random.default_rng(8)
    x0 = np.array([0.8, 330.0, 0.8])[None] * (1 + 0.02 * rng.uniform(-1, 1, (78000, 3)))
    acts = rng.uniform(-1, 1, (K, 1, 78000))
    outs, kernels = [], []
    for n, req in ((4096, "auto"), (20000, "auto"), (78000, "fast"), (78000, "auto")):
        env = _env(dict(params, x0=x0[:n]), n_envs=n)
        env.reset()
        with forced_kernel(req):
            out = env.rollout(torch.as_tensor(acts[:, :, :n].copy()))
        kernels.append(nat.lib().pcg_last_kernel().decode().split(" ")[0])
        outs.append({k: out[k][..., :m].cpu().numpy().astype(np.float64) for k in ("x", "obs", "reward", "done", "violated")})
        assert int(env.status.max()) == 0
    # 78 000 envs are more than 8 tiles per SM: the library runs the per-SM kernel in two waves of CTAs
    assert kernels == ["ws_rollout_kernel", "wsm_rollout_kernel", "fast_rollout_kernel", "wsm_rollout_kernel"], kernels
    for other in outs[1:]:
        np.testing.assert_allclose(other["x"], outs[0]["x"], rtol=2e-12, atol=0)
        np.testing.assert_allclose(other["obs"], outs[0]["obs"], rtol=0, atol=1e-10)
        ok = outs[0]["violated"] == other["violated"]           # a constraint value within rounding of 0 may flip
        assert ok.mean() > 0.9999
        np.testing.assert_allclose(other["reward"][ok], outs[0]["reward"][ok], rtol=1e-9, atol=1e-12)
        assert np.array_equal(other["done"], outs[0]["done"])
    assert np.array_equal(outs[1]["x"], outs[0]["x"])             # same integrator code in both warp-specialised kernels
    assert np.array_equal(outs[3]["x"], outs[0]["x"])


# ----------------------------------------------------------------------------- golden do-mpc trajectories through CUDA
@pytest.mark.parametrize("model", ["cstr", "four_tank", "multistage_extraction", "crystallization", "biofilm_reactor"])
def test_golden_trajectories_cuda(model, golden_dir):
    """Every stored transition x[:, i] = Phi(x[:, i-1], u[:, i]) of the reference's shipped do-mpc data
     is replayed as one env of a batch, one step each, physical-unit actions."""
    d = np.load(os.path.join(golden_dir, f"traj_{model}.npz"))
    x, u, dt = d["x"], d["u"], float(d["dt"])
    xin, uin, want = x[:, 1:-1], u[:, 2:], x[:, 2:]
    n = xin.shape[1]
    meta = orhs.MODELS[model]
    sp_key = meta["states"][0]
    params = {
        "model": model, "N": 4, "tsim": 4 * dt, "n_envs": n,
        "SP": {sp_key: [0.0] * 4},
        "x0": np.concatenate([xin.T, np.zeros((n, 1))], axis=1),
        "a_space": {"low": np.full(uin.shape[0], -1.0), "high": np.ones(uin.shape[0])},
        "o_space": {"low": -np.ones(x.shape[0] + 1), "high": np.ones(x.shape[0] + 1)},
        "normalise_a": False, "normalise_o": False,
        "integrator": {"scheme": "tsit5", "adaptive": True, "rtol": 1e-10, "atol": 1e-12},
    }
    env = _env(params)
    env.reset()
    obs, r, done, trunc, info = env.step(torch.as_tensor(uin.T.copy()))
    got = obs[:, : x.shape[0]].T.cpu().numpy()
    assert rel(got, want) < 5e-8, rel(got, want)
    assert int(env.status.max()) == 0


# ----------------------------------------------------------------------------- batch parity vs the vectorised oracle
BATCH = {
    "cstr": dict(x0=[0.8, 330.0], a=([295.0], [302.0]), dt=0.25, jitter=0.02),
    "first_order_system": dict(x0=[0.5], a=([0.0], [10.0]), dt=0.1, jitter=0.9),
    "four_tank": dict(x0=[0.141, 0.112, 0.072, 0.42], a=([0.1, 0.1], [10.0, 10.0]), dt=1000 / 60, jitter=0.1),
    "multistage_extraction": dict(x0=[0.55, 0.3, 0.45, 0.25, 0.4, 0.20, 0.35, 0.15, 0.25, 0.1], a=([5.0, 10.0], [500.0, 1000.0]), dt=1.0, jitter=0.05),
    # EXTENSION (BASELINE config 4's "10-stage, ~20-state" wording): ten stages of the same balances
    "multistage_extraction_10stage": dict(x0=[0.55, 0.3, 0.5, 0.28, 0.45, 0.25, 0.42, 0.22, 0.4, 0.2, 0.37, 0.18, 0.35, 0.15, 0.3, 0.13, 0.27, 0.11, 0.25, 0.1],
                                          a=([5.0, 10.0], [500.0, 1000.0]), dt=1.0, jitter=0.05),
    "crystallization": dict(x0=[1478.00986666666, 22995.8230590611, 1800863.24079725, 248516167.940593, 0.15861523304, 0.5, 15.0], a=([30.0], [40.0]), dt=1.0, jitter=0.01),   # T < 25 C blows the moments up
    "biofilm_reactor": dict(x0=[11.0, 5.3, 3.7, 480.0] * 4, a=([0, 1, 0.05, 0.05, 0.05], [10, 30, 1, 1, 1]), dt=2.0, jitter=0.02),
    # further registry models; input-less models take one ignored action entry
    "complex_cstr": dict(x0=[0.8, 0.1, 0.05, 330.0], a=([295.0], [302.0]), dt=0.25, jitter=0.02),
    "nonsmooth_control": dict(x0=[1.0, 0.5], a=([-1.0], [1.0]), dt=0.1, jitter=0.5),
    "cstr_series_recycle": dict(x0=[60.0, 310.0, 50.0, 310.0], a=([1e-5, 1e-5, 290.0, 290.0], [5e-5, 5e-5, 310.0, 310.0]), dt=5.0, jitter=0.02),
    "distillation_column": dict(x0=list(np.linspace(0.9, 0.05, 9)), a=([1.0, 150.0], [5.0, 300.0]), dt=0.5, jitter=0.05),
    "multistage_extraction_reactive": dict(x0=[1.0, 0.3, 1.5, 0.2] * 5, a=([5.0, 10.0], [500.0, 1000.0]), dt=0.02, jitter=0.05),
    "heat_exchanger": dict(x0=[360.0, 340.0, 320.0] * 8, a=([0.5, 0.5, 350.0, 280.0], [5.0, 5.0, 400.0, 320.0]), dt=0.5, jitter=0.02),
    "polymerisation_reactor": dict(x0=[330.0, 4.0, 0.3], a=([0.01, 300.0, 1.0, 0.1], [1.0, 350.0, 8.0, 1.0]), dt=0.05, jitter=0.01),
    "photobioreactor": dict(x0=[1.0, 150.0, 0.01], a=([120.0, 0.0], [400.0, 40.0]), dt=1.0, jitter=0.05),
    "invariant_batch": dict(x0=[1.0, 0.8, 0.1, 0.05], a=([0.0], [1.0]), dt=0.01, jitter=0.1),
    "batch": dict(x0=[1.0, 0.2, 0.05, 320.0], a=([280.0], [350.0]), dt=0.5, jitter=0.02),
    "coupled_oscillator": dict(x0=[0.5 + 0.3 * np.sin(i) for i in range(20)], a=([0.0], [1.0]), dt=0.1, jitter=0.1),
    "disease": dict(x0=[0.9, 0.1, 0.05], a=([0.0], [0.2]), dt=1.0, jitter=0.05),
    "hydraulic_tank": dict(x0=[1.0, 0.5], a=([-1.0], [1.0]), dt=0.1, jitter=0.2),
}
EXT_MODELS = ["complex_cstr", "nonsmooth_control", "cstr_series_recycle", "distillation_column",
              "multistage_extraction_reactive", "heat_exchanger", "polymerisation_reactor", "photobioreactor",
              "invariant_batch", "batch", "coupled_oscillator", "disease", "hydraulic_tank"]


def batch_setup(model, n, K, seed=0):
    b = BATCH[model]
    rng = np.random.default_rng(seed)
    nx = len(b["x0"])
    x0 = np.array(b["x0"])[None] * (1 + b["jitter"] * rng.uniform(-1, 1, (n, nx)))
    meta = orhs.MODELS[model]
    key = meta["states"][min(1, nx - 1)]
    sp = np.linspace(0.5, 1.5, K + 1) * b["x0"][meta["states"].index(key)]
    lo, hi = np.array(b["a"][0], float), np.array(b["a"][1], float)
    params = {
        "model": model, "N": K + 1, "tsim": (K + 1) * b["dt"], "n_envs": n,
        "SP": {key: list(sp)}, "r_scale": {key: 3.0},
        "x0": np.concatenate([x0, np.full((n, 1), sp[0])], axis=1),
        "a_space": {"low": lo, "high": hi},
        "o_space": {"low": np.concatenate([0.5 * np.array(b["x0"]), [0.0]]), "high": np.concatenate([1.5 * np.array(b["x0"]), [2 * sp.max()]])},
    }
    acts = rng.uniform(-1, 1, (K, len(lo), n))
    return params, x0, acts, key, sp


def oracle_batch(model, params, x0, acts, key, sp, n_sub):
    """Vectorised restatement of the step for SP-reward, noise-free, unconstrained batches."""
    lo, hi = params["a_space"]["low"], params["a_space"]["high"]
    dt = params["tsim"] / params["N"]
    meta = orhs.MODELS[model]
    si = meta["states"].index(key)
    x = x0.T.copy()
    xs, rs = [], []
    for k in range(acts.shape[0]):
        u = (acts[k] + 1) * ((hi - lo) / 2)[:, None] + lo[:, None]
        x = oint.advance_dense(model, x, u, dt, n_sub=n_sub)
        xs.append(x.copy())
        rs.append(-((x[si] - sp[k + 1]) ** 2) * 3.0)
    return np.stack(xs), np.stack(rs)


@pytest.mark.parametrize("model,n_sub_oracle,integ,tol", [
    ("cstr", 32, TIGHT, 1e-9),
    ("first_order_system", 32, TIGHT, 1e-9),
    ("four_tank", 256, TIGHT, 1e-8),                         # sqrt(h) is singular at h -> 0: slow convergence
    ("crystallization", 1024, dict(TIGHT, n_sub=200), 1e-7),
    ("biofilm_reactor", 1024, dict(TIGHT, n_sub=200), 1e-7),
    ("multistage_extraction", 2048, dict(TIGHT, n_sub=800), 5e-8),   # stiff: |lambda| dt up to 240
    ("cstr", 32, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}, 1e-6),
    ("four_tank", 256, {"scheme": "dopri5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}, 1e-6),
    ("multistage_extraction", 2048, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}, 1e-6),
    ("crystallization", 1024, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-9, "atol": 1e-12}, 1e-6),
    ("biofilm_reactor", 1024, {"scheme": "dopri5", "adaptive": True, "rtol": 1e-9, "atol": 1e-10}, 1e-6),
    ("cstr", 32, {"scheme": "rk4", "n_sub": 16}, 1e-8),
    # the stiff integrator (Rosenbrock RODAS4 with the analytic Jacobian) on the two models that provide one
    ("cstr", 32, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-8}, 1e-6),
    ("multistage_extraction", 2048, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-10}, 1e-6),
    # the labelled ten-stage extension: explicit pair and Rosenbrock
    ("multistage_extraction_10stage", 2048, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10}, 1e-6),
    ("multistage_extraction_10stage", 2048, {"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-10}, 1e-6),
] + [(m, 64, TIGHT, 1e-7 if m == "nonsmooth_control" else 1e-8)      # x2 crosses zero: relative error is loose there
     for m in EXT_MODELS if m != "multistage_extraction_reactive"] + [
    ("multistage_extraction_reactive", 512, dict(TIGHT, n_sub=200), 1e-7),      # fast gas-phase transients, tiny YC
    ("complex_cstr", 64, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}, 1e-6),
    ("heat_exchanger", 64, {"scheme": "dopri5", "adaptive": True, "rtol": 1e-8, "atol": 1e-8}, 1e-6),
    ("multistage_extraction_reactive", 512, {"scheme": "tsit5", "adaptive": True, "rtol": 1e-9, "atol": 1e-13}, 1e-6),
])
def test_batch_rollout_vs_oracle(model, n_sub_oracle, integ, tol):
    n, K = (2048 if n_sub_oracle <= 256 else 192), 12
    params, x0, acts, key, sp = batch_setup(model, n, K)
    env = _env(params, integrator=integ)
    env.reset()
    out = env.rollout(torch.as_tensor(acts))
    xs, rs = oracle_batch(model, params, x0, acts, key, sp, n_sub_oracle)
    got_x = out["x"].cpu().numpy()
    assert rel(got_x, xs) < tol, (model, rel(got_x, xs))
    np.testing.assert_allclose(out["reward"].cpu().numpy(), rs, rtol=max(tol * 1e3, 1e-6), atol=1e-9 * np.abs(rs).max())
    # observation = normalised [x | SP[t]]
    lo, hi = params["o_space"]["low"], params["o_space"]["high"]
    want_obs = 2 * (np.concatenate([xs, np.broadcast_to(sp[:K, None, None], (K, 1, n))], axis=1) - lo[None, :, None]) / (hi - lo)[None, :, None] - 1
    np.testing.assert_allclose(out["obs"].cpu().numpy(), want_obs, rtol=0, atol=max(tol * 10, 1e-9))
    assert int(env.status.max()) == 0
    assert out["done"][-1].all() and not out["done"][:-1].any()


def test_rodas4_follows_an_igniting_cstr_where_the_explicit_pairs_are_flagged():
    """Feed at 365 K (the reference's own disturbance range, pc-gym_paper/disturbance_showcase) with the coolant at its
    upper bound: the reactor ignites within a few intervals (T -> 450 K, the Jacobian's stiff eigenvalue times dt ~ 1e3).
    The reference integrates that with CVODES' BDF (integrator.py:163-182).  Here: scheme "rodas4" follows it -- against
    the oracle's own RODAS4 (same scheme, finite-difference Jacobian: 1e-6) and against the converged stiff solution
    (scipy Radau 1e-12: 1e-5 at rtol 1e-8) -- while a fixed explicit step is flagged in `status` instead of returning
    garbage silently."""
    N, n = 24, 64
    p = {"model": "cstr", "N": N, "tsim": N * 0.25, "n_envs": n, "SP": {"Ca": [0.85] * N},
         "o_space": {"low": np.array([0.0, 300.0, 0.8]), "high": np.array([1.0, 500.0, 0.9])},
         "a_space": {"low": np.array([295.0]), "high": np.array([302.0])},
         "x0": np.array([0.8, 330.0, 0.85]), "disturbances": {"Ti": np.full(N, 365.0)},
         "disturbance_bounds": {"low": np.array([330.0]), "high": np.array([370.0])}}
    acts = torch.linspace(0.6, 1.0, n, dtype=torch.float64)[None, None].repeat(N - 1, 1, 1)
    env = _env(p, integrator={"scheme": "rodas4", "rtol": 1e-8, "atol": 1e-8})
    env.reset()
    out = env.rollout(acts)
    torch.cuda.synchronize()
    assert int(env.status.max()) == 0
    got = out["x"].cpu().numpy()
    assert got[:, 1].max() > 440.0                              # it did ignite
    for i in (0, n // 2, n - 1):
        u = np.array([(acts[0, 0, i].item() + 1) * 3.5 + 295.0, 365.0, 1.0])
        xs, xt, h = np.array([0.8, 330.0]), np.array([0.8, 330.0]), None
        for k in range(N - 1):
            xs, h, _ = oint.advance_rodas4("cstr", xs, u, 0.25, rtol=1e-8, atol=1e-8, h0=h)
            xt = oint.advance_truth("cstr", xt, u, 0.25, method="Radau")
            assert rel(got[k, :, i], xs) < 1e-6, (i, k, rel(got[k, :, i], xs))
            assert rel(got[k, :, i], xt) < 1e-5, (i, k, rel(got[k, :, i], xt))
    steps = env.rhs_evaluations.double().mean().item() / 7 / (N - 1)
    assert 5 < steps < 80, steps                                # a few dozen steps per interval through the ignition
    bad = _env(p, integrator={"scheme": "tsit5", "adaptive": False, "n_sub": 1})
    bad.reset()
    bad.rollout(acts)
    assert int((bad.status != 0).sum()) == n                    # explicit, one step per interval: every env flagged
    with pytest.raises(ValueError):
        _env(dict(p, model="four_tank", disturbances=None, a_space={"low": np.array([0.1, 0.1]), "high": np.array([10.0, 10.0])},
                  x0=np.array([0.1, 0.1, 0.1, 0.1, 0.85]), SP={"h1": [0.85] * N},
                  o_space={"low": np.zeros(5), "high": np.ones(5)}), integrator={"scheme": "rodas4"})


def _oracle_fixed(model, params, x0, acts, scheme="tsit5", n_sub=1):
    """Noise-free trajectory [K, nx, n] by the oracle's own fixed-step scheme (same tableau, numpy arithmetic)."""
    lo, hi = np.asarray(params["a_space"]["low"], float), np.asarray(params["a_space"]["high"], float)
    dt = params["tsim"] / params["N"]
    x = x0.T.copy()
    xs = []
    for k in range(acts.shape[0]):
        u = (acts[k] + 1) * ((hi - lo) / 2)[:, None] + lo[:, None]
        x = oint.advance_fixed(model, x, u, dt, scheme=scheme, n_sub=n_sub)
        xs.append(x.copy())
    return np.stack(xs)


class forced_kernel:
    """pcg_set_kernel for the duration of a with-block (the request is per thread and sticky)."""
    def __init__(self, name):
        from pcgym_b200 import _native as nat
        self.nat, self.name = nat, name

    def __enter__(self):
        self.prev = self.nat.lib().pcg_set_kernel(self.nat.KERNELS[self.name])
        return self

    def __exit__(self, *a):
        self.nat.lib().pcg_set_kernel(self.prev)

    def last(self):
        return self.nat.lib().pcg_last_kernel().decode()


# BASELINE batch sizes: first_order 65 536, cstr 65 536, extraction 32 768; the others a quarter of that.
# The comparison is scheme against scheme (same tableau, one step per interval), so dt only has to keep that single
# explicit step well inside its stability region for random inputs: the documented dt of cstr / first_order, a
# fraction of it for the others (the documented intervals of those need several substeps: test_batch_rollout_vs_oracle).
MATRIX_DT = {"multistage_extraction": 0.004, "four_tank": 2.0, "crystallization": 0.25, "biofilm_reactor": 0.25}
MATRIX = [
    ("cstr", 65536, "wsm", "wsm_rollout_kernel"),
    ("cstr", 65536, "ws", "ws_rollout_kernel"),
    ("cstr", 65536, "fast", "fast_rollout_kernel (software-pipelined, 2 envs/thread)"),
    ("cstr", 65536, "fast_ept1", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("cstr", 65536, "generic", "rollout_kernel (generic, 2 envs/thread)"),
    ("cstr", 65535, "auto", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),       # odd batch: 64-bit rows
    ("cstr", 20000, "auto", "wsm_rollout_kernel"),                                           # partial last CTA
    ("first_order_system", 65536, "fast", "fast_rollout_kernel (software-pipelined, 2 envs/thread)"),
    ("first_order_system", 65536, "fast_ept1", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("first_order_system", 65536, "generic", "rollout_kernel (generic, 2 envs/thread)"),
    ("multistage_extraction", 32768, "fast", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("multistage_extraction", 32768, "generic", "rollout_kernel (generic, 1 env/thread)"),
    ("four_tank", 16384, "fast", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("four_tank", 16384, "generic", "rollout_kernel (generic, 1 env/thread)"),
    ("crystallization", 16384, "fast", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("crystallization", 16384, "generic", "rollout_kernel (generic, 1 env/thread)"),
    ("biofilm_reactor", 8192, "fast", "fast_rollout_kernel (software-pipelined, 1 env/thread)"),
    ("biofilm_reactor", 8192, "generic", "rollout_kernel (generic, 1 env/thread)"),
]


@pytest.mark.parametrize("model,n,request_,kernel", MATRIX)
def test_every_straight_line_kernel_against_the_oracle_same_scheme(model, n, request_, kernel):
    """Every kernel family x every north-star model at the BASELINE batch sizes, one Tsit5 step per control interval
    (the scheme the straight-line kernels are built for), two constraints on state 0, SP reward with penalty:
    x against the oracle's own Tsit5 step (same tableau, numpy arithmetic: 1e-11), reward / observation / flags
    against the reference formulae evaluated on the oracle's x.  pcg_last_kernel() proves which kernel ran."""
    K = 12
    params, x0, acts, key, sp = batch_setup(model, n, K)
    if model in MATRIX_DT:
        params["tsim"] = params["N"] * MATRIX_DT[model]
    x00 = BATCH[model]["x0"][0]
    xs = _oracle_fixed(model, params, x0, acts)
    # a band on state 0 that part of the population leaves: between the 15th and 85th percentile of the oracle's own values
    c_lo, c_hi = (float(v) for v in np.quantile(xs[:, 0], [0.15, 0.85]))
    params.update(constraints=lambda x, u: np.array([c_lo - x[0], x[0] - c_hi]), done_on_cons_vio=False, r_penalty=True,
                  constraint_units="physical", integrator=HEADLINE)
    with forced_kernel(request_) as fk:
        env = _env(params)
        env.reset()
        out = env.rollout(torch.as_tensor(acts))
        torch.cuda.synchronize()
        assert fk.last().startswith(kernel), fk.last()
    got_x = out["x"].cpu().numpy()
    per_state = np.array([rel(got_x[:, i], xs[:, i]) for i in range(xs.shape[1])])
    tol = np.full(xs.shape[1], 1e-11)
    if model == "crystallization":
        # dCV (model_classes.py:1313-1314) is a difference of products of the moments divided by 2 CV with
        # CV^2 = mu2 mu0 / mu1^2 - 1: two cancellations.  Perturbing the oracle's OWN state by one rounding error per
        # step moves CV by up to 1.2e-10 over these 12 steps (2 000 envs, numpy only), every other state by < 4e-14,
        # so for that one state "same scheme" can only mean agreement to the conditioning of the formula.
        tol[5] = 1e-7
    assert (per_state < tol).all(), (model, request_, per_state)
    meta = orhs.MODELS[model]
    si = meta["states"].index(key)
    g = np.stack([c_lo - xs[:, 0], xs[:, 0] - c_hi], axis=1)                       # [K, 2, n]
    viol = (g > 0).any(axis=1)
    sure = (np.abs(g) > 1e-9 * abs(x00)).all(axis=1)                                # a g within rounding of 0 may flip
    got_v = out["violated"].cpu().numpy().astype(bool)
    assert np.array_equal(got_v[sure], viol[sure]) and sure.mean() > 0.999 and viol.any() and not viol.all()
    want_r = -((xs[:, si] - sp[1:K + 1, None]) ** 2) * 3.0 - 1000.0 * viol
    np.testing.assert_allclose(out["reward"].cpu().numpy()[sure], want_r[sure], rtol=1e-9, atol=1e-9 * max(1.0, np.abs(want_r).max()))
    np.testing.assert_allclose(env.info["cons_info"][:, 1:].permute(1, 0, 2).cpu().numpy(), g, rtol=0, atol=1e-10 * x00)
    lo, hi = params["o_space"]["low"], params["o_space"]["high"]
    want_obs = 2 * (np.concatenate([xs, np.broadcast_to(sp[:K, None, None], (K, 1, n))], axis=1) - lo[None, :, None]) / (hi - lo)[None, :, None] - 1
    np.testing.assert_allclose(out["obs"].cpu().numpy(), want_obs, rtol=0, atol=1e-10)
    assert int(env.status.max()) == 0
    assert out["done"][-1].all() and not out["done"][:-1].any()


@pytest.mark.parametrize("tiles", [2, 3, 4, 5, 6, 7, 8])
def test_per_sm_kernel_every_tile_count_long_horizon(tiles):
    """The per-SM kernel at every tile count it accepts (2..8 tiles of 64 envs per SM), last tile partial, over a horizon
    that wraps its 8-slot rings a dozen times.  With few tiles per SM the integrator warps of different tiles share the
    four schedulers unevenly (3 tiles: 2 + 2 + 1 + 1 warps) and drift apart by a ring's depth -- the regime where a
    parity-only mbarrier wait a whole phase ahead stalled the kernel before the full_a gate (csrc/rollout_wsm.cuh).
    Same integrator arithmetic as the pair-per-CTA kernel up to the exponential's table form: states to 2e-12, and the
    run must terminate."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = 2 * ((sms * (64 * tiles - 10)) // 2)
    params = cases.cstr_dist_cons()
    params.pop("disturbances"); params.pop("disturbance_bounds")
    N = 100
    params.update(noise=True, noise_percentage=0.001, integrator=HEADLINE, seed=5, N=N, tsim=N * 0.25,
                  SP={"Ca": [0.85] * (N // 2) + [0.9] * (N - N // 2)})
    rng = np.random.default_rng(tiles)
    acts = torch.as_tensor(rng.uniform(-1, 1, (N - 1, 1, n))).cuda()
    outs = []
    for req, want in (("wsm", "wsm_rollout_kernel"), ("fast", "fast_rollout_kernel")):
        with forced_kernel(req) as fk:
            env = _env(params, n_envs=n)
            env.reset()
            out = env.rollout(acts)
            torch.cuda.synchronize()
            assert fk.last().startswith(want), (tiles, n, fk.last())
            assert int(env.status.max()) == 0
            outs.append({k: out[k].clone() for k in ("x", "obs", "reward", "done", "violated")})
    a, b = outs
    assert rel(a["x"].cpu().numpy(), b["x"].cpu().numpy()) < 2e-12
    np.testing.assert_allclose(a["obs"].cpu().numpy(), b["obs"].cpu().numpy(), rtol=0, atol=1e-10)
    assert torch.equal(a["done"], b["done"])
    assert (a["violated"] == b["violated"]).double().mean() > 0.9999


@pytest.mark.parametrize("per_env,n_sub,n", [(False, 1, 20000), (True, 1, 20000), (True, 2, 21846), (False, 3, 40000),
                                             ("none", 2, 65536)])
def test_general_per_sm_kernel_matches_the_generic_kernel(per_env, n_sub, n):
    """The GENERAL variant of the per-SM kernel (disturbance rows through the ring, observed disturbance slots,
    n_sub fixed substeps: what BASELINE config 5's cstr part needs) against the generic kernel on the same inputs:
    the reference's shared disturbance schedule (docs/guides/disturbances.md) and per-env tables, one to three substeps,
    constraints + penalty + noise on.  Same arithmetic up to the exponential's form in the integrator warps: states to
    2e-12, everything derived from them to rounding, identical flags away from rounding-level constraint values."""
    params = cases.cstr_dist_cons()
    N = params["N"]
    rng = np.random.default_rng(17)
    if per_env == "none":
        params.pop("disturbances"); params.pop("disturbance_bounds")
    elif per_env:
        params["disturbances"] = {"Ti": np.repeat(rng.uniform(340, 352, (n, (N + 3) // 4)), 4, axis=1)[:, :N],
                                  "Caf": rng.uniform(0.9, 1.1, (n, N))}
    else:
        params["disturbances"] = {"Ti": np.repeat(rng.uniform(340, 352, N // 4), 4), "Caf": rng.uniform(0.9, 1.1, N)}
    params.update(noise=True, noise_percentage=0.001, seed=9,
                  integrator={"scheme": "tsit5", "adaptive": False, "n_sub": n_sub})
    acts = torch.as_tensor(rng.uniform(-1, 1, (N - 1, 1, n))).cuda()
    outs = []
    for req, want in (("auto", "wsm_rollout_kernel (persistent CTA per SM, general"), ("generic", "rollout_kernel (generic")):
        with forced_kernel(req) as fk:
            env = _env(params, n_envs=n)
            env.reset()
            out = env.rollout(acts)
            torch.cuda.synchronize()
            assert fk.last().startswith(want), (req, fk.last())
            assert int(env.status.max()) == 0
            outs.append(({k: out[k].cpu().numpy().astype(np.float64) for k in ("x", "obs", "reward", "done", "violated")},
                         env.info["cons_info"].cpu().numpy(), env.rhs_evaluations.cpu().numpy()))
    (a, ga, ra), (b, gb, rb) = outs
    assert a["obs"].shape[1] == (3 if per_env == "none" else 5)              # [Ca, T, SP, Ti, Caf]
    assert rel(a["x"], b["x"]) < 2e-12
    np.testing.assert_allclose(a["obs"], b["obs"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(ga, gb, rtol=0, atol=1e-9)
    ok = a["violated"] == b["violated"]
    assert ok.mean() > 0.9999 and a["violated"].any()
    np.testing.assert_allclose(a["reward"][ok], b["reward"][ok], rtol=1e-9, atol=1e-12)
    assert np.array_equal(a["done"], b["done"]) and np.array_equal(ra, rb)


@pytest.mark.parametrize("model,n,scheme", [("multistage_extraction", 4099, "tsit5"), ("multistage_extraction", 32768, "dopri5"),
                                            ("multistage_extraction_10stage", 1000, "tsit5")])
def test_lane_split_kernel_matches_the_generic_adaptive_kernel(model, n, scheme):
    """The lane-split, lane-asynchronous kernel (one lane per stage of the column, csrc/rollout_lanes.cuh) is what the
    adaptive pairs run on for multistage_extraction: against the generic kernel (same controller; only the order of the
    additions in the error norm differs, so the two agree to the controller's tolerance, not bitwise), with two
    constraints that part of the population violates, penalty, measurement noise, batch sizes that leave a warp and a
    group partial.  A rollout cut in three reproduces the single launch bit for bit (carried step size, sticky done)."""
    K = 12
    params, x0, acts, key, sp = batch_setup(model, n, K)
    c_lo, c_hi = 0.9 * BATCH[model]["x0"][0], 1.02 * BATCH[model]["x0"][0]
    params.update(constraints=lambda x, u: np.array([c_lo - x[0], x[0] - c_hi]), done_on_cons_vio=False, r_penalty=True,
                  constraint_units="physical", noise=True, noise_percentage=0.002, seed=4,
                  integrator={"scheme": scheme, "adaptive": True, "rtol": 1e-8, "atol": 1e-10})
    a = torch.as_tensor(acts).cuda()
    outs = []
    for req, want in (("auto", "lanes_rollout_kernel"), ("generic", "rollout_kernel (generic")):
        with forced_kernel(req) as fk:
            env = _env(params)
            env.reset()
            out = env.rollout(a)
            torch.cuda.synchronize()
            assert fk.last().startswith(want), (req, fk.last())
            assert int(env.status.max()) == 0
            outs.append(({k: out[k].cpu().numpy().astype(np.float64) for k in ("x", "obs", "reward", "done", "violated")},
                         env.info["cons_info"].cpu().numpy(), env.rhs_evaluations.cpu().numpy().astype(np.float64)))
    (u, gu, ru), (v, gv, rv) = outs
    assert rel(u["x"], v["x"]) < 2e-7, rel(u["x"], v["x"])
    np.testing.assert_allclose(u["obs"], v["obs"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(gu, gv, rtol=0, atol=1e-7)
    ok = u["violated"] == v["violated"]
    assert ok.mean() > 0.999 and u["violated"].any() and not u["violated"].all()
    np.testing.assert_allclose(u["reward"][ok], v["reward"][ok], rtol=1e-5, atol=1e-8)
    assert np.array_equal(u["done"], v["done"])
    assert abs(ru.mean() / rv.mean() - 1) < 0.02                     # the same work, up to a few step decisions
    # against the converged solution (north-star rtol 1e-6)
    xs, _ = oracle_batch(model, params, x0[:96], acts[:, :, :96], key, sp, 2048)
    assert rel(u["x"][:, :, :96], xs) < 1e-6
    # split rollout == single rollout, bit for bit
    env = _env(params)
    env.reset()
    parts = [env.rollout(a[:1])["x"].clone(), env.rollout(a[1:5])["x"].clone(), env.rollout(a[5:])["x"].clone()]
    assert same(torch.cat(parts), torch.as_tensor(u["x"]).cuda())


def test_branch_free_math_against_correctly_rounded_values():
    """pcg_math_probe: the device helpers the right-hand sides use instead of the CUDA library's exp / log / sqrt /
    divide (csrc/fastmath.cuh), over the ranges the models use, against mpmath (relative error; one ulp is 1.1e-16 to
    2.2e-16; measured on a B200: exp 6.4e-16, Horner form 4.5e-16, table form 1.6e-15 at |x| = 40, log 2.2e-16, sqrt and
    the Newton reciprocal correctly rounded on every sample, third-order reciprocal 1.6e-16): bounds a little above those;
    NaN in -> NaN out, sqrt(0) = 0, sqrt(-1) = NaN, exp saturates instead of overflowing for |x| up to ~1e9 and is
    non-finite (never a finite wrong value) beyond."""
    import ctypes
    import mpmath as mp
    from pcgym_b200 import _native as nat
    mp.mp.dps = 50
    rng = np.random.default_rng(3)

    def run(op, x):
        xd = torch.as_tensor(x, dtype=torch.float64, device="cuda")
        yd = torch.empty_like(xd)
        nat.check(nat.lib().pcg_math_probe(nat.MATH_OPS[op], ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(yd.data_ptr()),
                                           xd.numel(), None))
        torch.cuda.synchronize()
        return yd.cpu().numpy()

    def relerr(got, ref_mp):
        ref = np.array([float(v) for v in ref_mp])
        return float(np.max(np.abs(got - ref) / np.abs(ref)))

    worst = {}
    x = np.concatenate([rng.uniform(-40, 40, 3000), rng.uniform(-1, 1, 1000)])
    ref_exp = [mp.exp(mp.mpf(float(v))) for v in x]
    for op in ("exp", "exp_horner"):
        worst[op] = relerr(run(op, x), ref_exp)
    # the table form reduces with ONE constant ln2/256: its error grows with |x| (~|x| 0.8e-16, fastmath.cuh); the cstr's
    # Arrhenius argument ln k0 - E/(R T) stays within +-12 for 250..800 K
    ref_arr = np.array([float(v) for v in ref_exp])
    rel_tab = np.abs(run("exp_tab", x) - ref_arr) / ref_arr
    worst["exp_tab"], worst["exp_tab_40"] = float(rel_tab[np.abs(x) <= 12].max()), float(rel_tab.max())
    xp = np.concatenate([np.exp(rng.uniform(-40, 40, 3000)), 1 + rng.uniform(-0.5, 0.5, 1000)])
    got, ref = run("log", xp), np.array([float(mp.log(mp.mpf(float(v)))) for v in xp])
    err = np.abs(got - ref)
    worst["log"] = float(np.max(np.where(err <= 4e-16, 0.0, err / np.abs(ref))))
    worst["sqrt"] = relerr(run("sqrt", xp), [mp.sqrt(mp.mpf(float(v))) for v in xp])
    xr = np.concatenate([xp, -xp[:500]])
    ref_rcp = [1 / mp.mpf(float(v)) for v in xr]
    worst["rcp"] = relerr(run("rcp", xr), ref_rcp)
    worst["rcp_o3"] = relerr(run("rcp_o3", xr), ref_rcp)
    print("worst relative errors:", {k: f"{v:.2e}" for k, v in worst.items()})
    bound = {"exp": 8e-16, "exp_horner": 8e-16, "exp_tab": 1.2e-15, "exp_tab_40": 4e-15, "log": 4e-16, "sqrt": 2.3e-16,
             "rcp": 2.3e-16, "rcp_o3": 4e-16}
    assert all(worst[k] <= bound[k] for k in bound), worst
    assert np.array_equal(run("log", -xp[:64]), got[:64])                          # log|x|
    edge = run("sqrt", np.array([0.0, -1.0, np.nan, 4.0]))
    assert edge[0] == 0.0 and np.isnan(edge[1]) and np.isnan(edge[2]) and edge[3] == 2.0
    for op in ("exp", "exp_horner", "exp_tab", "log", "rcp"):
        assert np.isnan(run(op, np.array([np.nan, 1.0]))[0]), op
    big = run("exp", np.array([1e4, -1e4, 1e9, -1e9]))                             # saturates at 2^+-1000 instead of inf / 0
    assert np.all(np.isfinite(big)) and big[0] > 1e300 and big[1] < 1e-300 and big[2] > 1e300 and big[3] < 1e-300
    far = run("exp", np.array([1e300, -1e300]))       # beyond ~1.5e9 the reduction is meaningless: never a FINITE wrong value
    assert not np.any(np.isfinite(far))


def test_fixed_step_probe_warns_when_the_interval_is_too_long():
    """A fixed step has no error estimate: the first reset() of a fixed-step env runs a short probe rollout against a
    tight adaptive run and warns when the scheme is more than 1e-6 away (round-1 verdict: the headline's single Tsit5
    step holds the bar at dt = 0.25 but nothing told a user who picked it at the reference's own dt = 0.433,
    pc-gym_paper/.../cstr_visualisation.py:11-12).  Quiet at 0.25, loud at 0.433, off with integrator={"check": False}."""
    import warnings
    import bench
    params, _ = bench.cstr_headline(n_envs=4096, device="cuda", seed=0)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        env = _env(params)
        env.reset()
    assert 0 < env.fixed_step_probe_error < 1e-6
    N = params["N"]
    slow = dict(params, tsim=N * 0.433)
    with pytest.warns(RuntimeWarning, match="away from the converged solution"):
        env = _env(slow)
        env.reset()
    assert env.fixed_step_probe_error > 1e-6
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        env = _env(dict(slow, integrator=dict(slow["integrator"], check=False)))
        env.reset()
        assert env.fixed_step_probe_error is None
        env = _env(dict(slow, integrator=dict(slow["integrator"], n_sub=2)))     # two substeps hold the bar again
        env.reset()
    assert env.fixed_step_probe_error < 1e-6


@pytest.mark.parametrize("scheme", ["tsit5", "dopri5"])
def test_adaptive_kernel_against_the_restated_controller(scheme):
    """make_env's default integrator (embedded pair + step-size controller) against its numpy restatement
    (oracle.integrate.advance_adaptive: same tableau, same controller, step carried between intervals): the kernel
    evaluates the growth factor in fp32 and exp / 1/T by its own routines, so the step sequences agree to ~1e-7 in h and
    the states far inside the tolerance both runs are held to; the right-hand-side counts agree env by env except where
    an accept decision falls within rounding of err = 1."""
    n, K = 64, 10
    params = cases.cstr_basic()
    params.update(noise=False, n_envs=n, integrator={"scheme": scheme, "adaptive": True, "rtol": 1e-8, "atol": 1e-8})
    rng = np.random.default_rng(2)
    x0 = np.array([0.8, 330.0, 0.85])[None] * (1 + 0.02 * rng.uniform(-1, 1, (n, 3)))
    params["x0"] = x0
    acts = rng.uniform(-1, 1, (K, 1, n))
    env = _env(params)
    env.reset()
    out = env.rollout(torch.as_tensor(acts))
    got_x = out["x"].cpu().numpy()
    got_n = env.rhs_evaluations.cpu().numpy()
    lo, hi = params["a_space"]["low"], params["a_space"]["high"]
    same_count = 0
    for i in range(8):
        x, h, total = x0[i, :2].copy(), None, 0
        for k in range(K):
            u = np.array([(acts[k, 0, i] + 1) * (hi[0] - lo[0]) / 2 + lo[0], 350.0, 1.0])
            x, h, nr = oint.advance_adaptive("cstr", x, u, 0.25, scheme=scheme, rtol=1e-8, atol=1e-8, h0=h)
            total += nr
            np.testing.assert_allclose(got_x[k, :, i], x, rtol=2e-8, atol=0, err_msg=f"env {i} interval {k}")
        same_count += int(total == int(got_n[i]))
        assert abs(total - int(got_n[i])) <= 12, (i, total, int(got_n[i]))      # at most two step decisions apart
    assert same_count >= 6, same_count
    assert int(env.status.max()) == 0


@pytest.mark.parametrize("n", [65536, 20000])
def test_default_adaptive_scheme_in_the_per_sm_kernel(n):
    """make_env's DEFAULT integrator (adaptive Tsit5, rtol = atol = 1e-8, the reference's jax path integrator.py:65-88)
    on the headline configuration runs in the integrator warps of the per-SM kernel: against the generic kernel's
    adaptive path (same controller, library-free exp instead of the table exp: same accepted steps, states to 1e-10),
    against the converged solution (north-star rtol 1e-6), flags / rewards as functions of its own state; the carried
    step sizes and RHS counts survive a split rollout."""
    import bench
    from pcgym_b200 import _native as nat
    params, acts = bench.cstr_headline(n_envs=n, device="cuda", seed=0)
    params.pop("integrator")                                   # the default: {"scheme": "tsit5", "adaptive": True, 1e-8}
    env = _env(params)
    assert env.integrator["adaptive"] and env.integrator["scheme"] == "tsit5"
    env.reset(episode=0)
    out = {k: v.clone() for k, v in env.rollout(acts).items()}
    torch.cuda.synchronize()
    assert nat.lib().pcg_last_kernel().decode().startswith("wsm_rollout_kernel (persistent CTA per SM, adaptive")
    assert int(env.status.max()) == 0
    rhs_wsm = env.rhs_evaluations.clone()
    with forced_kernel("generic") as fk:
        ref = _env(params)
        ref.reset(episode=0)
        want = ref.rollout(acts)
        torch.cuda.synchronize()
        assert fk.last().startswith("rollout_kernel (generic")
    # the two kernels evaluate exp / 1/T differently (table + third-order reciprocal against the library-free forms), so a
    # step whose error estimate sits on the accept / reject edge, or whose fp32 growth factor sits on a rounding edge, can
    # come out differently in a few of the 16 M steps: agreement is then at the level of the controller's own tolerance
    # (measured 1.3e-9), not of rounding
    assert rel(out["x"].cpu().numpy(), want["x"].cpu().numpy()) < 5e-8
    assert (out["violated"] != want["violated"]).double().mean().item() < 1e-6 and torch.equal(out["done"], want["done"])
    np.testing.assert_allclose(out["obs"].cpu().numpy(), want["obs"].cpu().numpy(), rtol=0, atol=1e-7)
    # same controller decisions: identical RHS evaluation counts in (nearly) every env
    same = (rhs_wsm == ref.rhs_evaluations).double().mean().item()
    assert same > 0.999, same
    x = params["x0"][:, :2].cpu().numpy().T.copy()
    a = acts.cpu().numpy()
    got = out["x"].cpu().numpy()
    worst = 0.0
    for k in range(a.shape[0]):
        x = oint.advance_dense("cstr", x, (a[k] + 1) * 3.5 + 295.0, 0.25, n_sub=16)
        worst = max(worst, rel(got[k], x))
    assert worst < 1e-6, worst
    # split rollout: h_carry / n_rhs / status are carried per env across launches
    env2 = _env(params)
    env2.reset(episode=0)
    o1 = {k: v.clone() for k, v in env2.rollout(acts[:40]).items()}
    o2 = env2.rollout(acts[40:])
    assert torch.equal(torch.cat([o1["x"], o2["x"]]), out["x"])
    assert torch.equal(env2.rhs_evaluations, rhs_wsm)


def test_headline_config_full_size_noise_on():
    """BASELINE config 3 EXACTLY as bench.py times it -- 65 536 CSTR envs x 99 intervals, Tsit5 x 1, Philox noise ON,
    constraints + penalty, the per-SM warp-specialised kernel -- against the oracle:
      * x: same scheme 1e-11, converged solution at the north-star rtol 1e-6;
      * observation: the reference formula (pcgym.py:453-493) on the oracle's x with the documented draw
        z = Phi^-1 of the Philox word (tests/noise_model.py, scipy ndtri);
      * reward / violated / cons_info / done: the reference formulae on the oracle's x."""
    import bench
    import noise_model as nm
    from pcgym_b200 import _native as nat
    n = 65536
    params, acts = bench.cstr_headline(n_envs=n, device="cuda", seed=0)
    env = _env(params)
    env.reset()
    out = env.rollout(acts)
    torch.cuda.synchronize()
    assert nat.lib().pcg_last_kernel().decode().startswith("wsm_rollout_kernel")
    assert env._desc.noise == nat.NOISE["philox"]
    x0 = params["x0"][:, :2].cpu().numpy()
    a = acts.cpu().numpy()
    K = a.shape[0]
    p1 = {"a_space": params["a_space"], "tsim": params["tsim"], "N": params["N"]}
    xs = _oracle_fixed("cstr", p1, x0, a)
    got_x = out["x"].cpu().numpy()
    assert rel(got_x, xs) < 1e-11, rel(got_x, xs)
    x = x0.T.copy()
    worst = 0.0
    for k in range(K):
        x = oint.advance_dense("cstr", x, (a[k] + 1) * 3.5 + 295.0, 0.25, n_sub=16)
        worst = max(worst, rel(got_x[k], x))
    assert worst < 1e-6, worst
    sp = np.array(params["SP"]["Ca"])
    lo, hi = params["o_space"]["low"], params["o_space"]["high"]
    gid = np.arange(n)
    got_obs, got_r = out["obs"].cpu().numpy(), out["reward"].cpu().numpy()
    got_v, got_g = out["violated"].cpu().numpy().astype(bool), env.info["cons_info"].cpu().numpy()
    for k in range(K):
        z = nm.normals(env._desc.seed, gid, k + 1, 2)
        noisy = xs[k] * (1 + z * 0.001)                                              # pcgym.py:458
        want = np.concatenate([noisy, np.full((1, n), sp[k])])                        # SP[t] in the observation
        want = 2 * (want - lo[:, None]) / (hi - lo)[:, None] - 1
        np.testing.assert_allclose(got_obs[k], want, rtol=0, atol=5e-10, err_msg=f"obs {k}")
        g = np.stack([319 - xs[k, 1], xs[k, 1] - 331])
        np.testing.assert_allclose(got_g[:, k + 1], g, rtol=0, atol=1e-8)
        viol = (g > 0).any(axis=0)
        sure = (np.abs(g) > 1e-7).all(axis=0)
        assert np.array_equal(got_v[k][sure], viol[sure])
        r = -((xs[k, 0] - sp[k + 1]) ** 2) * 1e3 - 1000.0 * viol                       # SP[t+1], noise-free state
        np.testing.assert_allclose(got_r[k][sure], r[sure], rtol=1e-8, atol=1e-10)
    assert out["done"][-1].all() and not out["done"][:-1].any()
    assert int(env.status.max()) == 0


# ----------------------------------------------------------------------------- noise
@pytest.mark.parametrize("n,id0", [(8, 1000), (7, 1001)])      # even batch: 2 envs/thread kernels; odd: 1 env/thread
def test_device_noise_is_philox_inverse_cdf(n, id0):
    """z recovered from the kernel's noisy observation equals the documented draw (csrc/philox.cuh): Philox4x32-10
    block of (env pair gid >> 1, t, state pair), word 2 (gid & 1) + (state & 1), z = Phi^-1((w + 1/2) 2^-32) -- the
    host model uses scipy's ndtri, not the kernel's table."""
    import noise_model as nm
    params = cases.cstr_basic()
    seed = 0x1234567_89ABCDEF
    env = _env(params, n_envs=n, seed=seed, env_id0=id0, normalise_o=False, integrator=HEADLINE)
    env.reset()
    a = torch.zeros(3, 1, n, dtype=torch.float64)
    out = env.rollout(a)
    pct = 0.001
    for k in range(3):
        x = out["x"][k].cpu().numpy()
        z = (out["obs"][k, :2].cpu().numpy() - x) / (x * pct)
        want = nm.normals(env._desc.seed, id0 + np.arange(n), k + 1, 2)
        np.testing.assert_allclose(z, want, rtol=0, atol=2e-9)                # (obs - x) / (x pct) loses ~7 digits


def test_normal_transform_accuracy_over_the_word_range():
    """pcg_normal_pairs: the fp64 inverse-CDF draw of csrc/philox.cuh against scipy's ndtri on edge words, both
    tails and 2^20 random words (3e-12); the fp32 Box-Muller evaluation to fp32 accuracy."""
    import noise_model as nm
    from pcgym_b200 import _native as nat
    rng = np.random.default_rng(5)
    edge = np.array([0, 1, 2, 3, 7, 8, 0x1fffffff, 0x20000000, 0x20000001, 0x3fffffff, 0x40000000, 0x5fffffff, 0x60000000,
                     0x7fffffff, 0x80000000, 0x80000001, 0x9fffffff, 0xa0000000, 0xbfffffff, 0xc0000000, 0xdfffffff,
                     0xe0000000, 0xfffffffe, 0xffffffff, 0x6a09e667, 0xb504f333, 0xb504f334], dtype=np.uint64)
    tails = np.concatenate([rng.integers(0, 1 << 14, 4096, dtype=np.uint64),
                            np.uint64(2 ** 32 - 1) - rng.integers(0, 1 << 14, 4096, dtype=np.uint64)])
    w0 = np.concatenate([np.repeat(edge, len(edge)), tails, rng.integers(0, 2 ** 32, 1 << 20, dtype=np.uint64)])
    w1 = np.concatenate([np.tile(edge, len(edge)), tails[::-1], rng.integers(0, 2 ** 32, 1 << 20, dtype=np.uint64)])
    n = w0.size
    dw0 = torch.from_numpy(w0.astype(np.uint32).view(np.int32)).cuda()
    dw1 = torch.from_numpy(w1.astype(np.uint32).view(np.int32)).cuda()
    for f32 in (0, 1):
        z0 = torch.empty(n, dtype=torch.float64, device="cuda")
        z1 = torch.empty_like(z0)
        nat.check(nat.lib().pcg_normal_pairs(dw0.data_ptr(), dw1.data_ptr(), z0.data_ptr(), z1.data_ptr(), n, f32,
                                             torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        if f32:                                                   # 24-bit uniforms, Box-Muller of the word pair
            u0f = ((w0 >> np.uint64(8)).astype(np.float64) + 0.5) * 2.0 ** -24
            u1f = ((w1 >> np.uint64(8)).astype(np.float64) + 0.5) * 2.0 ** -24
            radf = np.sqrt(-2 * np.log(u0f))
            ref0, ref1 = radf * np.cos(2 * np.pi * u1f), radf * np.sin(2 * np.pi * u1f)
            scale = np.maximum(1.0, radf)
            assert np.max(np.abs(z0.cpu().numpy() - ref0) / scale) < 1e-3
            assert np.max(np.abs(z1.cpu().numpy() - ref1) / scale) < 1e-3
        else:
            assert np.max(np.abs(z0.cpu().numpy() - nm.normal_from_word(w0))) < 3e-12
            assert np.max(np.abs(z1.cpu().numpy() - nm.normal_from_word(w1))) < 3e-12
            np.testing.assert_allclose(z0.cpu().numpy(), nm.icdf_table_eval(w0), rtol=0, atol=5e-15)   # host restatement (no fma there)


@pytest.mark.parametrize("mode", ["philox", "philox_f32"])
def test_noise_statistics_and_shard_invariance(mode):
    params = cases.cstr_basic()
    n, K = 32768, 8
    a = torch.zeros(K, 1, n, dtype=torch.float64)
    env = _env(params, n_envs=n, seed=11, noise_mode=mode, normalise_o=False, integrator=HEADLINE)
    env.reset()
    out = env.rollout(a)
    x = out["x"]
    z = ((out["obs"][:, :2] - x) / (x * 0.001)).cpu().numpy().ravel()
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.std() - 1) < 0.01
    assert abs((z ** 4).mean() - 3) < 0.1                      # kurtosis of a Gaussian
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.01
    # the second half of the batch run as its own shard (env_id0 offset) sees identical draws
    half = _env(params, n_envs=n // 2, env_id0=n // 2, seed=11, noise_mode=mode, normalise_o=False, integrator=HEADLINE)
    half.reset()
    out2 = half.rollout(a[:, :, n // 2:])
    assert torch.equal(out2["obs"], out["obs"][:, :, n // 2:])
    # different seed -> different draws
    other = _env(params, n_envs=n, seed=12, noise_mode=mode, normalise_o=False, integrator=HEADLINE)
    other.reset()
    assert not torch.equal(other.rollout(a)["obs"], out["obs"])


# ----------------------------------------------------------------------------- robustness
def test_biofilm_pole_sets_status_and_terminates():
    """Documented x0 runs into the Monod pole S2 = -K2: the kernel must finish, flag the
    affected envs and leave healthy envs untouched."""
    p = cases.biofilm()
    n = 256
    x0 = np.tile(np.array([11.0, 5.3, 3.7, 480.0] * 4 + [1.5]), (n, 1))
    x0[::2] = np.array([2, 0.1, 10, 0.1] * 4 + [1.5])          # every other env starts at the documented x0
    env = _env(p, n_envs=n, x0=x0, N=40, tsim=80, SP={"S2_A": [1.5] * 40},
               integrator={"scheme": "dopri5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10, "max_steps": 300})
    env.reset()
    out = env.rollout(torch.zeros(39, 5, n, dtype=torch.float64))
    torch.cuda.synchronize()
    st = env.status.cpu().numpy()
    assert (st[1::2] == 0).all()
    assert (st[0::2] != 0).all()
    assert torch.isfinite(out["x"][:, :, 1::2]).all()


def test_step_cap_flag():
    params, x0, acts, key, sp = batch_setup("multistage_extraction", 64, 4)
    env = _env(params, integrator={"scheme": "dopri5", "adaptive": True, "rtol": 1e-8, "atol": 1e-10, "max_steps": 3})
    env.reset()
    env.rollout(torch.as_tensor(acts))
    assert (env.status.cpu().numpy() & 2).all()


def test_per_env_setpoints_and_disturbances():
    p = cases.cstr_dist_cons()
    n, N = 16, p["N"]
    rng = np.random.default_rng(0)
    sp = rng.uniform(0.8, 0.9, (n, N))
    ti = rng.uniform(340, 360, (n, N))
    caf = np.asarray(p["disturbances"]["Caf"])
    acts = rng.uniform(-1, 1, (N - 1, 1, n))
    env = _env(p, n_envs=n, SP={"Ca": sp}, disturbances={"Ti": ti, "Caf": caf}, integrator=TIGHT)
    env.reset()
    out = env.rollout(torch.as_tensor(acts))
    x = np.tile(np.array([[0.8], [330.0]]), (1, n))
    for k in range(N - 1):
        u = np.stack([(acts[k, 0] + 1) * 3.5 + 295.0, ti[:, k + 1], np.full(n, caf[k + 1])])
        x = oint.advance_dense("cstr", x, u, p["tsim"] / N, n_sub=32)
        assert rel(out["x"][k].cpu().numpy(), x) < 1e-9
        np.testing.assert_allclose(out["reward"][k].cpu().numpy() + 1000.0, -((x[0] - sp[:, k + 1]) ** 2) * 1e3, rtol=1e-6, atol=1e-9)
        obs = out["obs"][k].cpu().numpy()
        np.testing.assert_allclose(obs[2], 2 * (sp[:, k] - 0.8) / 0.1 - 1, atol=1e-12)
        np.testing.assert_allclose(obs[3], 2 * (ti[:, k + 1] - 330.0) / 40.0 - 1, atol=1e-12)


def test_uncertainty_is_observation_only():
    p = cases.cstr_basic()
    p.update(noise=False, uncertainty_percentages={"k0": 0.1, "x0": [0.05, 0.0, 0.0]}, distribution="uniform",
             uncertainty_bounds={"low": np.array([6e10]), "high": np.array([8e10])})
    n = 512
    env = _env(p, n_envs=n, integrator=TIGHT)
    obs, _ = env.reset()
    assert obs.shape == (n, 4)
    k0 = (obs[:, 3].cpu().numpy() + 1) / 2 * 2e10 + 6e10
    assert k0.std() > 0 and np.all(np.abs(k0 / 7.2e10 - 1) <= 0.1 + 1e-12)
    ca0 = (obs[:, 0].cpu().numpy() + 1) / 2 * 0.3 + 0.7
    assert ca0.std() > 0 and np.all(np.abs(ca0 / 0.8 - 1) <= 0.05 + 1e-12)
    o1, *_ = env.step(torch.zeros(n, 1, dtype=torch.float64))
    np.testing.assert_allclose(o1[:, 3].cpu().numpy(), obs[:, 3].cpu().numpy())      # carried, never fed to the dynamics
    x = oint.advance_dense("cstr", np.stack([ca0, np.full(n, 330.0)]), np.full((1, n), 298.5), 0.25, n_sub=32)
    assert rel(env._x.cpu().numpy(), x) < 1e-9


def test_reference_known_answer_custom_model_through_cuda(golden_dir):
    """The reference's only numeric assertion on the integration path
    (tests/environment/test_make_env_custom_model.py:66-86: custom linear model, x0 = [1, 1], u = 0.5, dt = 0.1,
    obs ~= [1.21578082, 1.28403262] under np.isclose) run through the CUDA path: the custom model's affine RHS is
    read off the Python callable and integrated by the Affine<2,1> kernel."""
    from test_env_host import _Custom, _kat_params
    g = np.load(os.path.join(golden_dir, "kat_custom_model.npz"))
    env = _env(_kat_params(custom_model=_Custom(param1=1.5, param2=2.5)))
    obs0, _ = env.reset()
    assert isinstance(env.model, _Custom) and obs0.shape == (2,) and np.allclose(obs0, [1.0, 1.0])
    np.testing.assert_allclose(obs0, g["obs0"], rtol=0, atol=1e-15)
    obs, r, done, trunc, info = env.step(np.array([0.5]))
    assert obs.shape == (2,) and isinstance(r, float) and done is False and trunc is False
    assert np.all(np.isclose(obs, g["asserted"]))                            # the reference test's own check (rtol 1e-5)
    exact = np.array([(1 + 1 / 3) * np.exp(0.15) - 1 / 3, np.exp(0.25)])     # closed form of the linear ODE
    np.testing.assert_allclose(obs, exact, rtol=1e-9)
    np.testing.assert_allclose(obs, g["obs"], rtol=1e-9)                     # reference make_env + converged scipy
    np.testing.assert_allclose(r, float(g["reward"]), rtol=1e-8)
    # batched, 2 inputs, offset vector
    class Two:
        int_method = "casadi"
        def __call__(self, x, u):
            return np.array([-x[0] + 0.5 * x[2] + u[0] + 0.1, -2 * x[1] + u[1], x[0] - x[2] - 0.3])
        def info(self):
            return {"parameters": {}, "states": ["a", "b", "c"], "inputs": ["p", "q"], "disturbances": []}
    n, K = 64, 6
    rng = np.random.default_rng(2)
    x0 = rng.uniform(-1, 1, (n, 3))
    p = {"custom_model": Two(), "n_envs": n, "N": K + 1, "tsim": (K + 1) * 0.2, "SP": {"b": [0.0] * (K + 1)},
         "x0": np.concatenate([x0, np.zeros((n, 1))], axis=1), "normalise_a": False, "normalise_o": False,
         "a_space": {"low": -np.ones(2), "high": np.ones(2)}, "o_space": {"low": -np.ones(4), "high": np.ones(4)},
         "integrator": {"scheme": "tsit5", "adaptive": True, "rtol": 1e-11, "atol": 1e-13}}
    env = _env(p)
    env.reset()
    acts = rng.uniform(-1, 1, (K, 2, n))
    out = env.rollout(torch.as_tensor(acts))
    x = x0.T.copy()
    for k in range(K):
        x = oint.advance_dense(Two(), x, acts[k], 0.2, n_sub=64)
        np.testing.assert_allclose(out["x"][k].cpu().numpy(), x, rtol=1e-8, atol=1e-10, err_msg=str(k))


def test_per_env_parameters_reach_the_dynamics():
    """Row f4 (extension): `model_params` rows and, with uncertainty_in_dynamics, the sampled uncertain parameters
    drive each env's own RHS; the oracle integrates the same RHS with per-env parameter arrays."""
    n, K = 384, 10
    rng = np.random.default_rng(9)
    k0 = 7.2e10 * (1 + 0.03 * rng.uniform(-1, 1, n))       # larger spreads tip some envs into thermal runaway,
    ua = 5e4 * (1 + 0.05 * rng.uniform(-1, 1, n))          # where no integrator agrees with another to 1e-9
    p = cases.cstr_basic()
    p.update(noise=False, model_params={"k0": k0, "UA": ua})
    env = _env(p, n_envs=n, integrator=TIGHT)
    env.reset()
    acts = rng.uniform(-1, 1, (K, 1, n))
    out = env.rollout(torch.as_tensor(acts))
    par = orhs.default_params("cstr")
    par.update(k0=k0, UA=ua)
    x = np.tile(np.array([[0.8], [330.0]]), (1, n))
    for k in range(K):
        x = oint.advance_dense("cstr", x, (acts[k] + 1) * 3.5 + 295.0, 0.25, p=par, n_sub=32)
        assert rel(out["x"][k].cpu().numpy(), x) < 1e-9, k
    base = _env(dict(p, model_params=None), n_envs=n, integrator=TIGHT)
    base.reset()
    assert rel(base.rollout(torch.as_tensor(acts))["x"][-1].cpu().numpy(), x) > 1e-5      # the parameters mattered

    # an unsupplied disturbance channel falls back to the env's OWN parameter (Caf), a supplied one (Ti) wins
    caf = 1.0 + 0.05 * rng.uniform(-1, 1, n)
    ti = np.repeat(rng.uniform(345, 355, p["N"] // 4), 4)
    pd = cases.cstr_basic()
    pd.update(noise=False, model_params={"Caf": caf, "Ti": np.full(n, 999.0)}, disturbances={"Ti": ti},
              disturbance_bounds={"low": np.array([330.0]), "high": np.array([370.0])})
    envd = _env(pd, n_envs=n, integrator=TIGHT)
    envd.reset()
    outd = envd.rollout(torch.as_tensor(acts))
    x = np.tile(np.array([[0.8], [330.0]]), (1, n))
    for k in range(K):
        u = np.concatenate([(acts[k] + 1) * 3.5 + 295.0, np.full((1, n), ti[k + 1]), caf[None]])
        x = oint.advance_dense("cstr", x, u, 0.25, n_sub=32)
        assert rel(outd["x"][k].cpu().numpy(), x) < 1e-9, k

    # uncertainty_in_dynamics: the k0 reset() samples (and shows in the observation) is the k0 the env integrates with
    pu = cases.cstr_basic()
    pu.update(noise=False, uncertainty_percentages={"k0": 0.03}, distribution="uniform", uncertainty_in_dynamics=True,
              uncertainty_bounds={"low": np.array([6e10]), "high": np.array([8e10])})
    envu = _env(pu, n_envs=n, integrator=TIGHT)
    obs, _ = envu.reset()
    k0s = (obs[:, 3].cpu().numpy() + 1) / 2 * 2e10 + 6e10
    assert k0s.std() > 0
    outu = envu.rollout(torch.as_tensor(acts))
    par = orhs.default_params("cstr")
    par.update(k0=k0s)
    x = np.tile(np.array([[0.8], [330.0]]), (1, n))
    for k in range(K):
        x = oint.advance_dense("cstr", x, (acts[k] + 1) * 3.5 + 295.0, 0.25, p=par, n_sub=32)
    assert rel(outu["x"][-1].cpu().numpy(), x) < 1e-8


# ----------------------------------------------------------------------------- custom reward / declarative constraints
def test_fused_ocp_reward_matches_reference_function(golden_dir):
    """custom_reward={'kind': 'ocp_tracking'} is evaluated in the kernel; the fixture's rewards were produced by
    the reference's own sp_track_reward (pc-gym_paper/train_policies/cstr/custom_reward.py) inside its make_env."""
    g = np.load(os.path.join(golden_dir, "step_cstr_ocp.npz"))
    p = cases.cstr_ocp()
    p["custom_reward"] = {"kind": "ocp_tracking", "R": 0.1, "delta_u": True}
    K = g["actions"].shape[0]
    # step path, one env
    env = _env(p, noise_mode="external", integrator=TIGHT)
    env.reset()
    for k in range(K):
        obs, r, done, trunc, info = env.step(g["actions"][k].copy(), z=g["z"][k][None])
        np.testing.assert_allclose(r, g["reward"][k], rtol=1e-6, atol=1e-10, err_msg=f"reward {k}")
        np.testing.assert_allclose(obs, g["obs"][k], rtol=5e-7, atol=5e-8)
    # rollout path, batch of identical envs, one launch; u_prev carries over a reset like `env.u_prev` does
    n = 64
    env = _env(p, n_envs=n, noise_mode="external", integrator=TIGHT)
    env.reset()
    acts = torch.as_tensor(np.repeat(g["actions"].T[None].transpose(2, 1, 0), n, axis=2))       # [K, 1, n]
    zz = torch.as_tensor(np.repeat(g["z"][:, :, None], n, axis=2))
    out = env.rollout(acts, z=zz)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), np.repeat(g["reward"][:, None], n, axis=1), rtol=1e-6, atol=1e-10)
    env.reset()
    out = env.rollout(acts[:1], z=zz[:1])
    du = (g["actions"][0, 0] - g["actions"][-1, 0]) / 2.0                                        # normalised input change
    e = (out["x"][0, 0, 0].item() * (1 + 0.001 * g["z"][0, 0]) - 0.85) / 0.3
    np.testing.assert_allclose(out["reward"][0, 0].item(), -(e * e * 1e3 + 0.1 * du * du), rtol=1e-6)


@pytest.mark.parametrize("name,spec", [("four_tank_ocp", {"kind": "ocp_tracking", "R": 0.1}),
                                       ("cryst_ocp", {"kind": "cryst_tracking", "R": 0.01})])
def test_fused_paper_script_rewards_config5(name, spec, golden_dir):
    """BASELINE config 5's four_tank and crystallization shards with the FUSED form of their scripts' rewards
    (4tank_train.py:16-52, cryst_train.py:17-48): per-step rewards, observations and a_save against the fixtures the
    reference's make_env produced with the scripts' Python rewards (tests/golden/make_golden.py); step path and a
    64-env rollout in one launch."""
    g = np.load(os.path.join(golden_dir, f"step_{name}.npz"))
    p = cases.CASES[name]()
    p["custom_reward"] = spec
    K = g["actions"].shape[0]
    env = _env(p, noise_mode="external", integrator=dict(TIGHT, n_sub=48))
    env.reset()
    for k in range(K):
        obs, r, done, trunc, info = env.step(g["actions"][k].copy(), z=g["z"][k][None])
        np.testing.assert_allclose(r, g["reward"][k], rtol=1e-6, atol=1e-9, err_msg=f"{name} reward {k}")
        np.testing.assert_allclose(obs, g["obs"][k], rtol=5e-7, atol=5e-8, err_msg=f"{name} obs {k}")
    n = 64
    env = _env(p, n_envs=n, noise_mode="external", integrator=dict(TIGHT, n_sub=48))
    env.reset()
    acts = torch.as_tensor(np.repeat(g["actions"][:, :, None], n, axis=2))                      # [K, nu, n]
    zz = torch.as_tensor(np.repeat(g["z"][:, :, None], n, axis=2))
    out = env.rollout(acts, z=zz)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), np.repeat(g["reward"][:, None], n, axis=1), rtol=1e-6, atol=1e-9)
    assert int(env.status.max()) == 0


def test_python_callable_custom_reward_batched():
    p = cases.cstr_basic()
    n = 128
    def reward(env, obs, uk, violated):                      # torch-vectorised: [n, Nx], [n, Nu], [n]
        return -(obs[:, 0] - 0.9) ** 2 - 1e-3 * uk[:, 0] - violated.double()
    env = _env(p, n_envs=n, custom_reward=reward, noise=False, integrator=TIGHT)
    env.reset()
    a = torch.linspace(-1, 1, n, dtype=torch.float64, device="cuda")[:, None]
    obs, r, done, trunc, info = env.step(a)
    ca = env._x[0]
    uk = (a[:, 0] + 1) * 3.5 + 295.0
    np.testing.assert_allclose(r.cpu().numpy(), (-(ca - 0.9) ** 2 - 1e-3 * uk).cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(obs[:, 0].cpu().numpy(), (2 * (ca - 0.7) / 0.3 - 1).cpu().numpy(), rtol=1e-12)
    with pytest.raises(NotImplementedError):
        env.rollout(torch.zeros(2, 1, n, dtype=torch.float64))


@pytest.mark.parametrize("n", [1, 16])
def test_non_affine_constraints_unfused_path_matches_the_oracle_env(n):
    """constraints = a callable that is NOT affine in (x, u): evaluated unfused in step() on the returned state (with
    the reference's un-normalisation quirk, pcgym.py:597-608), cons_info / violated / penalty / done as the reference
    restated by oracle.env.OracleEnv; rollout() refuses it."""
    from oracle.env import OracleEnv
    p = cases.cstr_dist_cons()
    p.update(noise=False, done_on_cons_vio=True, r_penalty=True,
             constraints=lambda x, u: np.array([(x[1] - 4000.0) ** 2 - 3.0e7, u[0] * x[0] - 1328.0]))
    ref = OracleEnv(p, integrator="truth")
    ref.reset()
    env = _env(p, n_envs=n, integrator=TIGHT)
    env.reset()
    rng = np.random.default_rng(0)
    seen = set()
    for k in range(p["N"] - 1):
        a = rng.uniform(-1, 1, 1)
        o_ref, r_ref, d_ref, _, info_ref = ref.step(a)
        obs, r, done, trunc, info = env.step(a.copy() if n == 1 else torch.as_tensor(np.repeat(a[None], n, 0)))
        if n == 1:
            np.testing.assert_allclose(obs, o_ref, rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(r, r_ref, rtol=1e-6, atol=1e-9)
            assert done == d_ref
        else:
            np.testing.assert_allclose(obs.cpu().numpy(), np.repeat(o_ref[None], n, 0), rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(r.cpu().numpy(), np.full(n, r_ref), rtol=1e-6, atol=1e-9)
            assert bool(done.all()) == d_ref and bool(done.any()) == d_ref
        seen.add(bool(np.any(info_ref["cons_info"][:, ref.t, 0] > 0)))
        if d_ref:
            break
    ci = info["cons_info"].cpu().numpy()
    np.testing.assert_allclose(ci[:, : ref.t + 1, 0], info_ref["cons_info"][:, : ref.t + 1, 0], rtol=1e-6, atol=1e-3)
    assert seen == {False, True}                     # the run crosses from feasible to violated (and then ends)
    with pytest.raises(NotImplementedError):
        env.reset(); env.rollout(torch.zeros(2, 1, n, dtype=torch.float64))


def test_integration_md_stub_drives_pcg_step(tmp_path):
    """the `integration_method="b200"` stub a maintainer would add to the reference is
    executed AS WRITTEN in the document (minimal descriptor: nsp = 0, no noise / reward / constraints, output
    pointers NULL) and must return the converged ODE advance, with and without disturbance channels."""
    import re
    import types
    md = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    code = next(b for b in re.findall(r"```python\n(.*?)```", md, re.S) if "class b200_engine" in b)
    ns = {}
    exec(compile(code, "INTEGRATION.md#b200_engine", "exec"), ns)
    meta = orhs.MODELS["cstr"]
    model = types.SimpleNamespace(info=lambda: {"parameters": dict(meta["params"]), "states": meta["states"],
                                                "inputs": meta["inputs"], "disturbances": meta["disturbances"]})
    for dist in (False, True):
        env = types.SimpleNamespace(env_params={"model": "cstr"}, model=model, N=20, dt=0.25, t=0, disturbance_active=dist)
        eng = ns["b200_engine"](env)
        x = np.array([0.8, 330.0, 0.85])                       # state vector [x | SP slot]: only [:Nx_oracle] is read
        for t, u in enumerate(([299.0, 352.0, 1.03], [296.5, 345.0, 0.97], [301.0, 349.0, 1.0])):
            env.t = t
            uk = np.array(u if dist else u[:1])
            got = eng.step(x, uk)
            want = oint.advance_truth("cstr", x[:2], uk, 0.25)
            np.testing.assert_allclose(got, want, rtol=1e-7)
            x[:2] = got


def test_box_constraint_dict_api():
    """constraints={'T': [327, 321]}, cons_type={'T': ['<=', '>=']} (oracle.py:127-134, constraint_paper.py:33-35)."""
    p = cases.cstr_basic()
    p.update(noise=False, constraints={"T": [327.0, 321.0], "Tc": [301.0]}, cons_type={"T": ["<=", ">="], "Tc": ["<="]},
             done_on_cons_vio=False, r_penalty=True)
    n = 256
    env = _env(p, n_envs=n, integrator=HEADLINE)
    assert env.n_con == 3
    env.reset()
    acts = torch.rand(env.N - 1, 1, n, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 2 - 1
    out = env.rollout(acts)
    T_ = out["x"][:, 1]
    u = (acts[:, 0].cuda() + 1) * 3.5 + 295.0
    want = (T_ > 327.0) | (T_ < 321.0) | (u > 301.0)
    assert torch.equal(out["violated"].bool(), want) and want.any() and not want.all()
    g = env.info["cons_info"]
    np.testing.assert_allclose(g[0, 1:].cpu().numpy(), (T_ - 327.0).cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(g[2, 1:].cpu().numpy(), (u - 301.0).cpu().numpy(), rtol=0, atol=1e-12)
