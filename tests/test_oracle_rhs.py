"""oracle.rhs vs the reference's own model_classes.py (fixtures from tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import rhs as orhs

MODELS = ["cstr", "first_order_system", "multistage_extraction", "four_tank", "biofilm_reactor", "crystallization",
          "complex_cstr", "nonsmooth_control", "cstr_series_recycle", "distillation_column",
          "multistage_extraction_reactive", "heat_exchanger", "polymerisation_reactor", "photobioreactor",
          "invariant_batch", "batch", "coupled_oscillator", "disease", "hydraulic_tank"]


@pytest.fixture(scope="module")
def vec(golden_dir):
    return np.load(os.path.join(golden_dir, "rhs_vectors.npz"))


@pytest.mark.parametrize("model", MODELS)
def test_rhs_matches_reference(vec, model):
    x, u, dx = vec[f"{model}.x"], vec[f"{model}.u"], vec[f"{model}.dx"]
    got = orhs.rhs(model, x.T, u.T).T              # vectorised over the 64 sample points
    np.testing.assert_allclose(got, dx, rtol=2e-14, atol=1e-300)
    # and one point at a time (the shape the env oracle uses)
    one = orhs.rhs(model, x[3], u[3])
    np.testing.assert_allclose(one, dx[3], rtol=2e-14, atol=0)


@pytest.mark.parametrize("model", ["cstr", "multistage_extraction", "complex_cstr"])
def test_rhs_with_disturbance_inputs(vec, model):
    """u longer than the manipulated inputs overrides Ti/Caf resp. X0/Y6 (model_classes.py:48-51, 382-385)."""
    x, ud, dx = vec[f"{model}.x"], vec[f"{model}.ud"], vec[f"{model}.dx_d"]
    np.testing.assert_allclose(orhs.rhs(model, x.T, ud.T).T, dx, rtol=2e-14, atol=0)


def test_survey_anchor_values():
    """Anchors quoted in SURVEY 8(c) (reference RHS evaluated in the survey container)."""
    np.testing.assert_allclose(orhs.rhs("first_order_system", [0.25], [5.0]), [9.5], rtol=1e-15)
    np.testing.assert_allclose(orhs.rhs("cstr", [0.8, 330.0], [300.0]),
                               [0.02419205660909973, -5.981601801066908], rtol=1e-13)


def test_metadata_matches_reference_info():
    assert orhs.MODELS["cstr"]["disturbances"] == ["Ti", "Caf"]
    assert orhs.MODELS["crystallization"]["states"] == ["Mu0", "Mu1", "Mu2", "Mu3", "Conc", "CV", "Ln"]
    assert len(orhs.MODELS["biofilm_reactor"]["states"]) == 16
    assert orhs.MODELS["biofilm_reactor"]["states"][:5] == ["S1_1", "S2_1", "S3_1", "O_1", "S1_2"]
    assert orhs.MODELS["four_tank"]["disturbances"] == ["None"]


def test_registry_is_complete():
    """Every key of the reference's model_mapping (src/pcgym/pcgym.py:128-148) has an oracle RHS."""
    keys = ["cstr", "complex_cstr", "first_order_system", "nonsmooth_control", "multistage_extraction",
            "cstr_series_recycle", "distillation_column", "multistage_extraction_reactive", "four_tank",
            "heat_exchanger", "biofilm_reactor", "polymerisation_reactor", "photobioreactor", "crystallization",
            "invariant_batch", "batch", "coupled_oscillator", "disease", "hydraulic_tank"]
    extension = ["multistage_extraction_10stage"]          # not a reference model: see test below
    assert sorted(keys + extension) == sorted(orhs.RHS) == sorted(orhs.MODELS)


def test_ten_stage_extension_shares_the_five_stage_balances():
    """The ten-stage column is an EXTENSION (BASELINE config 4's wording; no reference model of that size).  What pins
    it: (a) it runs the very code path that the reference-pinned five-stage RHS runs; (b) a ten-stage column whose
    stages 6..10 sit exactly at the gas-feed equilibrium hands stage 5 the same inlet the five-stage column gets, so its
    first ten derivatives equal the five-stage RHS; (c) the column's overall mass balance,
    sum_i (Vl dX_i + Vg dY_i) = L (X0 - X_N) + G (Y_{N+1} - Y_1), for random states."""
    rng = np.random.default_rng(0)
    p5, p10 = orhs.default_params("multistage_extraction"), orhs.default_params("multistage_extraction_10stage")
    u = np.array([120.0, 300.0])
    x5 = rng.uniform(0.05, 0.6, 10)
    x10 = np.concatenate([x5, np.tile([0.0, p5["Y6"]], 5)])     # Y6..Y10 = the five-stage gas feed
    d5 = orhs.rhs("multistage_extraction", x5, u)
    d10 = orhs.rhs("multistage_extraction_10stage", x10, u)
    np.testing.assert_allclose(d10[:10], d5, rtol=1e-14, atol=1e-15)
    for name, p, n, yfeed in (("multistage_extraction", p5, 5, p5["Y6"]), ("multistage_extraction_10stage", p10, 10, p10["Y11"])):
        x = rng.uniform(0.05, 0.6, 2 * n)
        d = orhs.rhs(name, x, u)
        total = p["Vl"] * d[0::2].sum() + p["Vg"] * d[1::2].sum()
        np.testing.assert_allclose(total, u[0] * (p["X0"] - x[2 * n - 2]) + u[1] * (yfeed - x[1]), rtol=1e-12)
