#!/usr/bin/env python
"""Fixture for pcgym_b200.evaluation_metrics: runs the REFERENCE's reproducibility_metric
(/root/reference/src/pcgym/evaluation_metrics.py, imported verbatim -- it only needs numpy) on seeded rollout-shaped
data and stores inputs + outputs in tests/golden/metrics_fixture.npz.  Build container only.

    python tests/golden/make_golden_metrics.py
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/src/pcgym/evaluation_metrics.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(42)
N, reps = 12, 7
data = {"pi_a": {"r": rng.normal(size=(1, N, reps)), "x": rng.normal(size=(3, N, reps)),
                 "u": rng.normal(size=(1, N, reps)), "g": rng.normal(size=(2, N, 1, reps))},
        "pi_b": {"r": rng.normal(size=(1, N, reps)), "x": rng.normal(size=(3, N, reps)),
                 "u": rng.normal(size=(1, N, reps)), "g": rng.normal(size=(2, N, 1, reps))}}
out = {}
for pol, comps in data.items():
    for c, v in comps.items():
        out[f"in/{pol}/{c}"] = v
for disp, perf, w in [("std", "mean", -1.0), ("std", "median", 2.0)]:
    m = ref.reproducibility_metric(disp, perf, w)
    for name, res in (("scal", m.scalarised_performance(data, None)), ("perf", m.policy_performance_metric(data, None)),
                      ("disp", m.policy_dispersion_metric(data, None))):
        for pol in res:
            for c in res[pol]:
                out[f"{disp}_{perf}/{name}/{pol}/{c}"] = res[pol][c]
# MAD: the reference expression only broadcasts as intended for a single time index (evaluation_metrics.py:113-116)
single = {"pi_a": {"r": rng.normal(size=(1, 1, reps))}}
out["in_single/pi_a/r"] = single["pi_a"]["r"]
m = ref.reproducibility_metric("mad", "median", -1.0)
out["mad_median/scal/pi_a/r"] = m.scalarised_performance(single, "r")["pi_a"]["r"]
np.savez_compressed(os.path.join(HERE, "metrics_fixture.npz"), **out)
print("wrote metrics_fixture.npz with", len(out), "arrays")
