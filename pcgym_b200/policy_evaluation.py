"""`policy_eval`: the reference's closed-loop rollout collector (src/pcgym/policy_evaluation.py:6-197) with the
`reps` axis mapped onto the environment batch.

The reference runs `reps` episodes one after the other, each a Python loop of N-1 `policy.predict` /
`env.step` calls on ONE environment (policy_evaluation.py:71-130, 155-197).  Repetitions differ only through
the measurement noise, the sampled uncertain parameters / x0 and a stochastic policy, so here they are the
`n_envs = reps` lanes of one batched environment: an episode costs N-1 kernel launches for ALL repetitions, and
for device policies the whole episode is captured once into a CUDA graph and replayed.

Same constructor, same `rollout` / `get_rollouts` results: the dictionary
    {policy_name: {"r": (1, N, reps), "x": (Nx, N, reps), "u": (nu, N, reps)[, "g": (n_con, N, 1, reps)]}}
of float64 numpy arrays (`as_torch=True` keeps device tensors).  Out of scope here: the
NMPC "oracle" rows (`oracle=True`, do-mpc/IPOPT) and matplotlib plotting (`plot_data`) raise.

Policies, in order of preference
  * `policy.predict_batch(obs) -> actions` or a `torch.nn.Module` / plain callable: called with the device tensor
    obs [reps, Nx] (float64), returns actions [reps, nu] on the device -- no host round trip, graph-capturable;
  * the reference's Stable-Baselines3 interface `policy.predict(obs, deterministic=True) -> (action, state)`:
    called with a float64 numpy batch [reps, Nx] (one device->host->device round trip per step).
"""
from __future__ import annotations

import numpy as np
import torch


def _is_device_policy(p) -> bool:
    return hasattr(p, "predict_batch") or isinstance(p, torch.nn.Module) or (callable(p) and not hasattr(p, "predict"))


class policy_eval:
    def __init__(self, make_env, policies: dict, reps: int, env_params: dict, oracle: bool = False,
                 MPC_params=False, cons_viol: bool = False, save_fig: bool = False, cuda_graph: bool = False,
                 as_torch: bool = False):
        self.make_env = make_env
        self.env_params = env_params
        self.policies = policies
        self.n_pi = len(policies)
        self.reps = int(reps)
        self.oracle = oracle
        self.cons_viol = cons_viol
        self.save_fig = save_fig
        self.MPC_params = MPC_params
        self.cuda_graph = cuda_graph
        self.as_torch = as_torch
        if oracle:
            raise NotImplementedError("oracle=True needs the do-mpc/IPOPT NMPC of pcgym.oracle, which is outside the "
                                      "B200 hot path (DESIGN.md section 9)")
        p = dict(env_params)
        p["n_envs"] = self.reps                      # reps == env batch
        self.env = make_env(p)
        self._graphs = {}
        self._o_lo = self._o_hi = self._a_lo = self._a_hi = None

    # ------------------------------------------------------------------ one policy, all repetitions
    def _act(self, policy, obs):
        """obs: device tensor [reps, Nx] -> actions: device tensor [reps, nu] (float64)."""
        env = self.env
        if hasattr(policy, "predict_batch"):
            a = policy.predict_batch(obs)
        elif _is_device_policy(policy):
            a = policy(obs)
        else:                                        # Stable-Baselines3 style, policy_evaluation.py:101
            o = obs.detach().cpu().numpy()
            if self.reps == 1:
                a, _ = policy.predict(o[0], deterministic=True)
                a = np.asarray(a, dtype=np.float64).reshape(1, -1)
            else:
                a, _ = policy.predict(o, deterministic=True)
                a = np.asarray(a, dtype=np.float64).reshape(self.reps, -1)
            a = torch.from_numpy(np.ascontiguousarray(a))
        return torch.as_tensor(a).to(device=env.device, dtype=torch.float64).reshape(self.reps, -1)

    def _unnorm_obs(self, o):
        """(o + 1)(hi - lo)/2 + lo with observation_space_base, applied unconditionally as the reference does
        (policy_evaluation.py:91-93, 108-110)."""
        lo, hi = self._o_lo, self._o_hi
        return (o + 1) * (hi - lo) / 2 + lo

    def _unnorm_act(self, a):
        return (a + 1) * (self._a_hi - self._a_lo) / 2 + self._a_lo                # policy_evaluation.py:103-107

    def _episode(self, policy, o, bufs):
        """N-1 x (predict, step) + the trailing predict (policy_evaluation.py:100-130); writes into `bufs`."""
        env = self.env
        rew, states, actions = bufs
        partial = bool(getattr(env, "partial_observation", None))
        for i in range(env.N - 1):
            a = self._act(policy, o)
            o, r, term, trunc, info = env.step(a)
            if o.ndim == 1:                                          # reps == 1 with a host policy never gets here
                o = o[None]
            actions[:, i] = self._unnorm_act(a).T
            states[:, i + 1] = self._unnorm_obs(info["obs"] if partial else o).T
            rew[0, i + 1] = r
        a = self._act(policy, o)
        actions[:, env.N - 1] = self._unnorm_act(a).T

    def rollout(self, policy_i):
        """All `reps` repetitions of one policy.  Returns (total_reward [N, reps], s_rollout [Nx, N, reps],
        actions [nu, N, reps], cons_info [n_con, N, 1, reps]) -- the reference's per-repetition tuple with the
        repetition index as a trailing axis -- as device tensors."""
        env = self.env
        dev, f64, R = env.device, torch.float64, self.reps
        nu = int(np.asarray(env.env_params["a_space"]["low"]).reshape(-1).shape[0])
        o, info = env.reset()
        if not isinstance(o, torch.Tensor):
            o = torch.as_tensor(np.asarray(o, dtype=np.float64), device=dev)[None]
        if self._o_lo is None:      # created once: a captured episode keeps reading these buffers
            self._o_lo = torch.as_tensor(np.asarray(env.observation_space_base.low, dtype=np.float64), device=dev)
            self._o_hi = torch.as_tensor(np.asarray(env.observation_space_base.high, dtype=np.float64), device=dev)
            self._a_lo = torch.as_tensor(np.asarray(env.env_params["a_space"]["low"], dtype=np.float64).reshape(-1), device=dev)
            self._a_hi = torch.as_tensor(np.asarray(env.env_params["a_space"]["high"], dtype=np.float64).reshape(-1), device=dev)
        rew = torch.zeros(1, env.N, R, dtype=f64, device=dev)
        states = torch.zeros(env.Nx, env.N, R, dtype=f64, device=dev)
        actions = torch.zeros(nu, env.N, R, dtype=f64, device=dev)
        rew[0, 0] = float(info["r_init"])
        partial = bool(getattr(env, "partial_observation", None))
        states[:, 0] = self._unnorm_obs(info["obs"] if partial else o).T           # policy_evaluation.py:90-98
        if self.cuda_graph and _is_device_policy(policy_i):
            self._episode_graphed(policy_i, o, (rew, states, actions))
        else:
            self._episode(policy_i, o, (rew, states, actions))
        if env.constraint_active:
            cons_info = info["cons_info"].reshape(env.n_con, env.N, 1, R).clone()
        else:
            cons_info = torch.zeros(1, env.N, 1, R, dtype=f64, device=dev)         # policy_evaluation.py:121-124
        return rew[0], states, actions, cons_info

    # ------------------------------------------------------------------ CUDA-graph replay of a whole episode
    def _episode_graphed(self, policy, o, bufs):
        """Device policies only: the N-1 (policy, step kernel) pairs of one episode are captured once per policy
        and replayed; the time index of every launch is part of the captured kernel arguments, the buffers are
        the environment's own (reset() rewrites them in place)."""
        env = self.env
        key = id(policy)
        g = self._graphs.get(key)
        if g is None:
            # warm-up episode on a side stream (allocator, lazy initialisation), then reset and capture.
            # reset() is deterministic in (seed, reset argument), so the state the capture starts from is the
            # state every later reset() re-creates in the same buffers.
            s = torch.cuda.Stream(device=env.device)
            s.wait_stream(torch.cuda.current_stream(env.device))
            o_static = o.clone()
            with torch.cuda.stream(s):
                self._episode(policy, o_static, tuple(b.clone() for b in bufs))
            torch.cuda.current_stream(env.device).wait_stream(s)
            torch.cuda.synchronize(env.device)
            # same episode index as the warm-up: the noise key is a by-value kernel argument, so the captured episode
            # replays THIS episode's measurement noise every time (a fresh key needs a fresh capture)
            env.reset(episode=env.episode)
            static_bufs = tuple(b.clone() for b in bufs)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._episode(policy, o_static, static_bufs)
            env.t, env.done = 0, False              # capture records launches, it does not run them
            g = self._graphs[key] = (graph, o_static, static_bufs)
        graph, o_static, static_bufs = g
        o_static.copy_(o)
        for dst, src in zip(static_bufs, bufs):
            dst.copy_(src)                          # row 0 (initial observation, r_init)
        graph.replay()
        env.t = env.N - 1
        for dst, src in zip(bufs, static_bufs):
            dst.copy_(src)

    # ------------------------------------------------------------------ all policies (policy_evaluation.py:155-197)
    def get_rollouts(self) -> dict:
        data = {}
        for pi_name, pi_i in self.policies.items():
            rew, states, actions, cons_info = self.rollout(pi_i)
            entry = {"r": rew[None], "x": states, "u": actions}
            if self.env.constraint_active:
                entry["g"] = cons_info
            if not self.as_torch:
                entry = {k: v.detach().cpu().numpy() for k, v in entry.items()}
            data[pi_name] = entry
        self.data = data
        return data

    def oracle_reward_fn(self, x, u):
        raise NotImplementedError("oracle rows are out of scope (DESIGN.md section 9)")

    def plot_data(self, data, reward_dist=False):
        raise NotImplementedError("plotting (matplotlib) is outside the B200 hot path; the arrays in `data` have "
                                  "the reference's shapes and can be passed to pcgym's own plot_data")
