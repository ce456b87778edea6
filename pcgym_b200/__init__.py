"""pcgym_b200: B200-native batched drop-in for pc-gym's `make_env.step` hot path.

    from pcgym_b200 import make_env          # same env_params dict as pcgym.make_env, plus n_envs
    env = make_env({... , "n_envs": 65536})
    obs, info = env.reset()
    obs, r, done, trunc, info = env.step(actions)      # one fused CUDA launch for all envs
    traj = env.rollout(actions_K)                       # K control intervals in one launch
    mb = MixedBatch({"cstr": p1, "four_tank": p2})      # sub-batches of different models advanced by one call

The numerics live in libpcgym_b200.so (hand-written sm_100a kernels behind the C ABI of
include/pcgym_b200.h); importing this package without the built library raises.
"""
from . import _native
from .env import make_env
from .mixed import MixedBatch
from .evaluation_metrics import reproducibility_metric
from .models import DeviceModel, model_mapping
from .policy_evaluation import policy_eval

_native.lib()   # no CPU fallback: fail at import time if the CUDA library is missing

__all__ = ["make_env", "MixedBatch", "reproducibility_metric", "policy_eval", "DeviceModel", "model_mapping"]
__version__ = "0.1.0"
