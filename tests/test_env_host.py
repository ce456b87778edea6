"""Host-side logic of pcgym_b200.make_env (no GPU): configuration parsing mirrors the reference
(checked against the oracle env, which is itself pinned to the reference), descriptor contents,
constraint probing, error behaviour."""
import numpy as np
import pytest

import cases
from oracle.env import OracleEnv
from pcgym_b200 import _native as nat
from pcgym_b200 import make_env

ATTRS = ["N", "tsim", "dt", "Nx", "Nx_oracle", "Nu", "Nd", "Nd_model", "normalise_a", "normalise_o", "a_delta",
         "constraint_active", "r_penalty", "disturbance_active", "custom_reward", "reward", "maximise_reward"]


@pytest.mark.parametrize("name", list(cases.CASES))
def test_attributes_match_reference_semantics(name):
    p = cases.CASES[name]()
    env, ref = make_env(p), OracleEnv(p)
    for a in ATTRS:
        assert getattr(env, a) == getattr(ref, a), a
    assert env.action_space.shape == ref.action_space.shape
    assert env.observation_space.shape == ref.observation_space.shape
    np.testing.assert_allclose(env.observation_space_base.low, ref.observation_space_base.low)
    np.testing.assert_allclose(env.observation_space_base.high, ref.observation_space_base.high)
    assert env.model.info()["states"] == ref.model.info()["states"]
    assert env.model.info()["disturbances"] == ref.model.info()["disturbances"]
    assert env.model.info()["parameters"] == ref.model.info()["parameters"]
    if env.constraint_active:
        assert env.n_con == ref.n_con and env.done_on_constraint == ref.done_on_constraint


def test_descriptor_contents_cstr():
    env = make_env(cases.cstr_dist_cons())
    d = env._desc
    assert (d.model, d.nu_a, d.nd, d.nsp, d.nd_obs, d.n_con) == (0, 1, 2, 1, 2, 2)
    assert list(d.d_src)[:2] == [0, 1] and list(d.d_obs)[:2] == [0, 1]
    assert d.flags & nat.F_NORMALISE_A and d.flags & nat.F_NORMALISE_O and d.flags & nat.F_R_PENALTY
    assert not d.flags & nat.F_DONE_ON_VIOLATION
    assert d.a_scale[0] == 3.5 and d.a_off[0] == 298.5
    assert d.sp_state[0] == 0 and d.r_scale[0] == 1e3 and d.sp_init[0] == 0.8
    assert d.dt == 0.25 and d.N == 40 and d.reward == nat.REWARD["sp"]
    # reference quirk folded in: g sees (T+1)*25+300 (pcgym.py:602-608)
    assert d.con_gx[0][1] == -25.0 and d.con_h[0] == 319 - 325 and d.con_mask_x[0] == 0b10
    phys = make_env(dict(cases.cstr_dist_cons(), constraint_units="physical"))._desc
    assert phys.con_gx[0][1] == -1.0 and phys.con_h[0] == 319.0 and phys.con_gx[1][1] == 1.0 and phys.con_h[1] == -331.0


def test_descriptor_one_disturbance_and_partial_obs():
    env = make_env(cases.cstr_one_dist())
    d = env._desc
    assert (d.nd, d.nd_obs) == (2, 1)
    assert list(d.d_src)[:2] == [-1, 0]                    # Ti from the model default, Caf from the table
    assert d.d_default[0] == 350.0 and d.d_obs[0] == 1
    assert d.obs_zero_mask == 0b10                         # T is not observed


def test_descriptor_noise_dict_and_batch_reward():
    d = make_env(cases.cstr_nonorm_cons())._desc
    assert list(d.noise_pct)[:2] == [0.0, 0.002] and d.noise == nat.NOISE["philox"]
    assert d.flags & nat.F_DONE_ON_VIOLATION and not d.flags & nat.F_NORMALISE_A
    assert d.con_gu[2][0] == 1.0 and d.con_mask_u[2] == 1
    b = make_env(cases.cstr_batch_reward())._desc
    assert b.reward == nat.REWARD["batch"] and list(b.batch_w)[:2] == [-10.0, -1.0] and b.nsp == 0


def test_a_delta_descriptor():
    d = make_env(cases.crystallization())._desc
    assert d.flags & nat.F_A_DELTA and d.a_act_lo[0] == 10.0 and d.a_act_hi[0] == 40.0
    assert d.a_scale[0] == 1.0 and d.a_off[0] == 0.0
    assert d.nsp == 2 and list(d.sp_state)[:2] == [5, 6]
    # a_delta without normalise_a is dead code in the reference (pcgym.py:376)
    p = cases.crystallization(); p["normalise_a"] = False
    assert not make_env(p)._desc.flags & nat.F_A_DELTA


def test_integrator_options():
    p = cases.first_order()
    assert make_env(p).integrator["scheme"] == "tsit5" and make_env(p)._desc.flags & nat.F_ADAPTIVE
    d = make_env(dict(p, integrator={"scheme": "rk4", "n_sub": 4}))._desc
    assert d.scheme == 0 and d.n_sub == 4 and not d.flags & nat.F_ADAPTIVE
    for im in ("casadi", "jax", "b200"):
        make_env(dict(p, integration_method=im))
    with pytest.raises(AssertionError):
        make_env(dict(p, integration_method="euler"))
    with pytest.raises(ValueError):
        make_env(dict(p, integrator={"scheme": "bdf"}))


def test_non_affine_constraints_take_the_unfused_path():
    """A callable that is not affine in (x, u) cannot be fused: make_env accepts it like the reference
    (pcgym.py:571-572), learns n_con from one call (:124), keeps the kernel's descriptor free of constraints and
    evaluates the callable in step(); rollout() refuses it."""
    p = cases.cstr_dist_cons()
    p["constraints"] = lambda x, u: np.array([x[1] ** 2 - 1e5, u[0] * x[0] - 400.0])
    env = make_env(p)
    assert env.constraint_active and env._host_constraints and env.n_con == 2
    assert env._desc.n_con == 0 and not (env._desc.flags & (nat.F_R_PENALTY | nat.F_DONE_ON_VIOLATION))
    affine = make_env(cases.cstr_dist_cons())
    assert not affine._host_constraints and affine._desc.n_con == 2


class _Custom:
    """The reference's example custom model (tests/environment/test_make_env_custom_model.py:7-26)."""

    def __init__(self, param1=1.5, param2=2.5, square=False):
        self.int_method, self.param1, self.param2, self.square = "casadi", param1, param2, square

    def __call__(self, x, u):
        return np.array([self.param1 * (x[0] ** 2 if self.square else x[0]) + u[0], self.param2 * x[1]])

    def info(self):
        return {"parameters": {"param1": self.param1, "param2": self.param2}, "states": ["x1", "x2"],
                "inputs": ["u1"], "disturbances": []}


def _kat_params(**kw):
    p = {"custom_model": _Custom(), "a_space": {"low": np.array([-1]), "high": np.array([1])},
         "o_space": {"low": np.array([-1, -1]), "high": np.array([1, 1])}, "SP": {"x2": [2] * 100}, "N": 100,
         "tsim": 10, "x0": np.array([1.0, 1.0])}
    p.update(kw)
    return p


def test_custom_model_is_probed_for_affinity():
    """env_params["custom_model"] (pcgym.py:150-153): env.model stays the user's object, int_method is set on it, the
    affine map is read off the callable; x0 / o_space without setpoint slots give an observation of the states only."""
    m = _Custom()
    env = make_env(_kat_params(custom_model=m))
    assert isinstance(env.model, _Custom) and env.model.int_method == "b200"     # a deep copy, as in the reference (pcgym.py:42)
    np.testing.assert_array_equal(env._dmodel.A, [[1.5, 0.0], [0.0, 2.5]])
    np.testing.assert_array_equal(env._dmodel.B, [[1.0], [0.0]])
    np.testing.assert_array_equal(env._dmodel.c, [0.0, 0.0])
    d = env._desc
    assert d.model == nat.AFFINE_BASE + 2 and list(d.theta)[:8] == [1.5, 0.0, 0.0, 2.5, 1.0, 0.0, 0.0, 0.0]
    assert (env.Nx_oracle, env.Nx, env._n_obs) == (2, 3, 2) and d.flags & nat.F_SP_HIDDEN and d.nsp == 1
    assert env.observation_space.shape == (2,)
    with pytest.raises(ValueError):
        make_env(_kat_params(integrator={"scheme": "dopri5"}))


def test_errors_like_reference():
    with pytest.raises(ValueError):
        make_env([1, 2])
    with pytest.raises(ValueError):
        make_env(dict(cases.first_order(), model="nope"))
    with pytest.raises(ValueError):                      # a_space length must match the model's inputs
        make_env(dict(cases.first_order(), model="heat_exchanger"))
    with pytest.raises(NotImplementedError):             # a non-affine Python RHS cannot run in the kernel
        make_env(dict(_kat_params(), custom_model=_Custom(square=True)))
    with pytest.raises(KeyError):                        # four_tank's only disturbance key is "None" (pcgym.py:400)
        make_env(dict(cases.four_tank(), disturbances={"q1": np.ones(24)},
                      disturbance_bounds={"low": np.array([0.0]), "high": np.array([2.0])}))
    with pytest.raises(IndexError):
        make_env(dict(cases.first_order(), SP={"x": [5.0] * 3}))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    env = make_env(cases.first_order())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset()
    with pytest.raises(RuntimeError):
        env.step(np.array([0.0]))
    with pytest.raises(NotImplementedError):
        env.model(np.zeros(1), np.zeros(1))


def test_per_env_tables():
    p = cases.cstr_basic()
    n = 4
    sp = np.tile(np.linspace(0.8, 0.9, p["N"]), (n, 1))
    env = make_env(dict(p, n_envs=n, SP={"Ca": sp}, x0=np.tile(p["x0"], (n, 1))))
    assert env._sp_per_env and env._sp_host.shape == (p["N"], 1, n)
    assert env._x0_host.shape == (3, n)
    with pytest.raises(ValueError):
        make_env(dict(p, n_envs=n, x0=np.zeros((3, 3))))


def test_custom_reward_forms():
    p = cases.cstr_ocp()
    env = make_env(p)                                              # Python callable: unfused path
    assert env.custom_reward and env._host_reward and env._desc.reward == nat.REWARD["none"]
    assert not env._desc.flags & nat.F_NORMALISE_O and env.normalise_o is True      # kernel emits physical obs
    fused = make_env(dict(p, custom_reward={"kind": "ocp_tracking", "R": 0.25, "delta_u": False}))
    d = fused._desc
    assert d.reward == nat.REWARD["ocp"] and d.ocp_R == 0.25 and d.ocp_mode == 0 and not fused._host_reward
    assert abs(d.ocp_inv_span[0] - 1 / 0.3) < 1e-12 and d.flags & nat.F_NORMALISE_O
    assert make_env(dict(p, custom_reward={"kind": "ocp_tracking"}))._desc.ocp_mode == 2
    with pytest.raises(ValueError):
        make_env(dict(p, custom_reward={"kind": "nope"}))
    with pytest.raises(ValueError):
        make_env(dict(p, custom_reward=3.0))


def test_box_constraint_dict_descriptor():
    p = cases.cstr_basic()
    p.update(constraints={"T": [327.0, 321.0], "Tc": [301.0]}, cons_type={"T": ["<=", ">="], "Tc": ["<="]},
             done_on_cons_vio=True, r_penalty=False)
    d = make_env(p)._desc
    assert d.n_con == 3 and d.flags & nat.F_DONE_ON_VIOLATION
    assert (d.con_gx[0][1], d.con_h[0]) == (1.0, -327.0) and (d.con_gx[1][1], d.con_h[1]) == (-1.0, 321.0)
    assert (d.con_gu[2][0], d.con_h[2]) == (1.0, -301.0) and d.con_gx[2][1] == 0.0
    with pytest.raises(KeyError):
        make_env(dict(p, cons_type=None))
    with pytest.raises(ValueError):
        make_env(dict(p, constraints={"Q": [1.0]}, cons_type={"Q": ["<="]}))


def test_every_registry_key_of_the_reference_constructs():
    """src/pcgym/pcgym.py:128-148: all 19 keys resolve to a device model whose dimensions the library confirms (plus the
    one labelled extension, the ten-stage extraction column)."""
    from oracle import rhs as orhs
    from pcgym_b200 import models
    extensions = {"multistage_extraction_10stage"}
    assert sorted(models.model_mapping()) == sorted(orhs.MODELS) and len(set(orhs.MODELS) - extensions) == 19
    for name, meta in orhs.MODELS.items():
        m = models.get(name)
        info = m.info()
        assert info["states"] == meta["states"] and info["inputs"] == meta["inputs"], name
        assert {k: float(v) for k, v in info["parameters"].items()} == {k: float(v) for k, v in meta["params"].items()}, name
        nx, nu, nd, npar = nat.model_dims(m.model_id)
        assert (nx, nu, npar) == (len(meta["states"]), max(len(meta["inputs"]), 1), len(meta["params"])), name
    assert "disturbances" not in models.get("heat_exchanger").info()           # model_classes.py:1036-1041
    assert models.get("photobioreactor").info()["uncertainties"] == []         # model_classes.py:497-505
    env = make_env({"model": "heat_exchanger", "N": 10, "tsim": 5.0, "SP": {"Tt8": [350.0] * 10},
                    "x0": np.array([360.0, 340.0, 320.0] * 8 + [350.0]),
                    "o_space": {"low": np.full(25, 250.0), "high": np.full(25, 450.0)},
                    "a_space": {"low": np.array([0.5, 0.5, 350.0, 280.0]), "high": np.array([5.0, 5.0, 400.0, 320.0])}})
    assert (env.Nx_oracle, env.Nx, env.Nu) == (24, 25, 4) and env._desc.model == 11
    with pytest.raises(NotImplementedError):
        models.DeviceModel("coupled_oscillator", N=12)


def test_affine_probe_recovers_the_matrices_and_rejects_oversize():
    from pcgym_b200 import models

    class M:
        int_method = "casadi"
        A = np.array([[-1.0, 0.0, 0.5], [0.0, -2.0, 0.0], [1.0, 0.0, -1.0]])
        B = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
        c = np.array([0.1, 0.0, -0.3])

        def __call__(self, x, u):
            return self.A @ x + self.B @ u + self.c

        def info(self):
            return {"parameters": {}, "states": ["a", "b", "c"], "inputs": ["p", "q"], "disturbances": []}
    dm = models.AffineCustomModel(M())
    np.testing.assert_allclose(dm.A, M.A, atol=1e-15)
    np.testing.assert_allclose(dm.B, M.B, atol=1e-15)
    np.testing.assert_allclose(dm.c, M.c, atol=1e-15)
    assert dm.model_id == nat.AFFINE_BASE + 2 * nat.AFFINE_MAX_NU + 1 and len(dm.theta()) == 9 + 6 + 3
    assert nat.model_dims(dm.model_id)[:2] == (3, 2)

    class Big(M):
        def info(self):
            return {"parameters": {}, "states": list("abcde"), "inputs": ["p"], "disturbances": []}
    with pytest.raises(NotImplementedError):
        models.AffineCustomModel(Big())


def test_math_tables_are_reproducible():
    """csrc/exp2tab.inc (2^(j/256), 60-digit decimal arithmetic) and csrc/icdftab.inc (inverse normal CDF segments) are
    exactly what their generators under tools/ write; spot values are checked against their definitions."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for gen, inc in (("gen_exp2tab.py", "exp2tab.inc"), ("gen_icdf_tab.py", "icdftab.inc")):
        want = open(os.path.join(root, "pcgym_b200", "csrc", inc)).read()
        src = open(os.path.join(root, "tools", gen)).read()
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "tools")); os.makedirs(os.path.join(tmp, "pcgym_b200", "csrc"))
            open(os.path.join(tmp, "tools", gen), "w").write(src)
            subprocess.run([sys.executable, os.path.join(tmp, "tools", gen)], check=True, capture_output=True)
            got = open(os.path.join(tmp, "pcgym_b200", "csrc", inc)).read()
        assert got == want, inc
    vals = [float(v) for ln in open(os.path.join(root, "pcgym_b200", "csrc", "exp2tab.inc")) if not ln.startswith("//")
            for v in ln.strip().rstrip(",").split(",")]
    assert len(vals) == 256 and vals[0] == 1.0 and abs(vals[128] - 2 ** 0.5) < 3e-16 and abs(vals[255] - 2 ** (255 / 256)) < 3e-16


def test_noise_transform_host_model_matches_the_inverse_normal_cdf():
    """tests/noise_model.icdf_table_eval restates the device algorithm on the committed table: within 3e-12 of
    scipy's ndtri over edge words, both tails and 2^21 random words; the Philox model passes the Random123 vectors."""
    import noise_model as nm
    rng = np.random.default_rng(1)
    edge = np.array([0, 1, 2, 3, 7, 8, 15, 16, 0x7fffffff, 0x80000000, 0x80000001, 0xffffffff, 0xfffffffe, 0x3fffffff,
                     0x40000000, 0x1fffffff, 0x20000000, 0x6a09e667, 0xb504f333], dtype=np.uint64)
    w = np.concatenate([edge, rng.integers(0, 2 ** 32, 1 << 21, dtype=np.uint64), rng.integers(0, 4096, 4096, dtype=np.uint64),
                        np.uint64(2 ** 32 - 1) - rng.integers(0, 4096, 4096, dtype=np.uint64)])
    z, ref = nm.icdf_table_eval(w), nm.normal_from_word(w)
    assert np.max(np.abs(z - ref)) < 3e-12
    assert np.all(np.diff(nm.icdf_table_eval(np.sort(w))) >= 0)                 # monotone across every segment boundary
    assert abs(ref[:2].min() + 6.33795775) < 1e-7 and abs(ref[11] - 6.33795775) < 1e-7
    assert [int(v) for v in nm.philox4x32_10(0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(v) for v in nm.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0x299f31d0a4093822)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
