"""Mixed-model batches (BASELINE config 5): ONE call advances sub-batches of different registry models.

The reference has no notion of a batch at all (one `make_env` = one environment, `pcgym.py:31-500`); a mixed
population there is a Python list of envs stepped in a loop.  Here each model family is one `make_env` over its
share of the population (its own descriptor, kernels and output buffers -- the state dimensions differ, so the
struct-of-arrays rows cannot be shared), and `MixedBatch` drives them as a unit:

    mb = MixedBatch({"cstr": cstr_params, "four_tank": ft_params, "crystallization": cr_params})
    obs0 = mb.reset()
    out = mb.rollout({"cstr": a_cstr, "four_tank": a_ft, "crystallization": a_cr})     # one call, concurrent launches

* every part's launches go to its own CUDA stream, forked from and joined to the caller's stream, so the kernels of
  the parts overlap on the device (they are different kernels with different per-SM footprints) and the call is a
  single node for the caller: after `rollout()` returns, work enqueued on the caller's stream sees every result;
* sharding (`rank`, `world`): EACH part's envs are split evenly (`sharding.mixed_plan`), so every GPU carries the
  same mix -- the parts cost different amounts per env-step and a split by model would leave GPUs idle.  There is
  no exchange step and therefore no collective;
* `rollout_host()` is the same for pinned host buffers (each part pipelines its own copies, `make_env.rollout_host`).
"""
from __future__ import annotations

import torch

from . import sharding
from .env import make_env


class MixedBatch:
    def __init__(self, parts: dict, rank: int = 0, world: int = 1, device=None):
        """parts: name -> env_params (as for `make_env`, with the GLOBAL `n_envs` of that part).  With world > 1 each
        part is cut to this rank's shard (`sharding.shard_params`: per-env x0 / SP / disturbance tensors sliced,
        `env_id0` offset so the noise streams equal the single-process run)."""
        if not parts:
            raise ValueError("MixedBatch needs at least one part")
        self.names = list(parts)
        self.plan = sharding.mixed_plan({k: int(p.get("n_envs", 1)) for k, p in parts.items()}, rank, world)
        self.envs = {}
        for name, p in parts.items():
            p = dict(p)
            if device is not None:
                p["device"] = device
            self.envs[name] = make_env(sharding.shard_params(p, rank, world) if world > 1 else p)
        devs = {str(e.device) for e in self.envs.values()}
        if len(devs) != 1:
            raise ValueError(f"all parts of a MixedBatch live on one device, got {sorted(devs)}")
        self.device = next(iter(self.envs.values())).device
        self._streams = None

    # ------------------------------------------------------------------ plumbing
    @property
    def n_envs(self) -> int:
        return sum(e.n_envs for e in self.envs.values())

    def _fork(self):
        if self._streams is None:
            self._streams = {k: torch.cuda.Stream(device=self.device) for k in self.names}
        main = torch.cuda.current_stream(self.device)
        for s in self._streams.values():
            s.wait_stream(main)
        return main

    def _join(self, main):
        for s in self._streams.values():
            main.wait_stream(s)

    def _each(self, fn):
        """fn(name, env) for every part on the part's stream; the caller's stream waits for all of them."""
        main = self._fork()
        out = {}
        for k in self.names:
            with torch.cuda.stream(self._streams[k]):
                out[k] = fn(k, self.envs[k])
        self._join(main)
        return out

    # ------------------------------------------------------------------ the env surface, per part
    def reset(self, seed: int = 0, **kw):
        """name -> initial observation [n_part, n_obs] (make_env.reset of every part)."""
        return {k: e.reset(seed=seed, **kw)[0] for k, e in self.envs.items()}

    def rewind(self, t: int = 0, episode=None, to_x0: bool = False):
        self._each(lambda k, e: e.rewind(t, episode, to_x0))

    def rollout(self, actions: dict, want_x: bool = True, copy: bool = False):
        """actions: name -> [K_part, nu_a, n_part] (the parts may have different horizons).  One call; the parts'
        kernels run concurrently.  Returns name -> the part's rollout dict (obs, reward, done, violated, x)."""
        missing = [k for k in self.names if k not in actions]
        if missing:
            raise KeyError(f"no actions for parts {missing}")
        return self._each(lambda k, e: e.rollout(actions[k], want_x=want_x, copy=copy))

    def step(self, actions: dict):
        """One control interval of every part: name -> (obs, reward, done, truncated, info)."""
        return self._each(lambda k, e: e.step(actions[k]))

    def rollout_host(self, actions: dict, out: dict | None = None, chunks: int = 4, wait: bool = True, outputs=None):
        """`rollout` with pinned HOST buffers (name -> [K, nu_a, n] float64): every part pipelines its own
        host->device / kernel / device->host slices on its own streams (make_env.rollout_host)."""
        res = {k: self.envs[k].rollout_host(actions[k], out=None if out is None else out.get(k), chunks=chunks,
                                            wait=False, outputs=outputs) for k in self.names}
        if wait:
            self.host_sync()
        return res

    def host_sync(self):
        for e in self.envs.values():
            e.host_sync()

    @property
    def status(self):
        """name -> per-env integrator status bits (0 = healthy)."""
        return {k: e.status for k, e in self.envs.items()}

    def status_nonzero(self) -> int:
        return int(sum(int((e.status != 0).sum()) for e in self.envs.values()))

    def env_steps(self, horizons: dict) -> int:
        """env-steps one rollout with these per-part horizons advances."""
        return sum(self.envs[k].n_envs * int(horizons[k]) for k in self.names)
