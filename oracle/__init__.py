"""CPU oracle for the pc-gym `make_env.step` hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference algorithm
(`/root/reference/src/pcgym/{pcgym,integrator,model_classes}.py`).  It exists so
that the CUDA path can be checked against it; nothing under `pcgym_b200/` may
import it.  The only legitimate callers are `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity pin (see DESIGN.md "Oracle"):
  * `oracle.rhs`  is checked against RHS values produced by the reference's own
    `model_classes.py` executed verbatim in the build container
    (`tests/golden/rhs_vectors.npz`, generator `tests/golden/make_golden.py`).
  * `oracle.integrate` is checked against the golden do-mpc trajectories the
    reference ships (`pc-gym_paper/train_policies/*/data.npy`, extracted to
    `tests/golden/traj_*.npz`), the reference's known-answer test
    (`tests/environment/test_make_env_custom_model.py:85-86`) and a closed form.
  * `oracle.env`  is checked against step-by-step outputs of the reference's
    own `make_env` class run verbatim (only the CasADi/CVODES call replaced by a
    converged scipy integrator, because CasADi is not installable here):
    `tests/golden/step_*.npz`.
"""
