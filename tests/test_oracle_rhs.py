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
    assert sorted(keys) == sorted(orhs.RHS) == sorted(orhs.MODELS)
