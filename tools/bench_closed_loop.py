#!/usr/bin/env python
"""Closed-loop throughput of policy_eval (row f1): a small device MLP policy in the loop, reps = env batch, one
(policy, step-kernel) pair per control interval -- eager launches against CUDA-graph replay of the whole episode.

    python tools/bench_closed_loop.py [--reps 65536] [--episodes 5]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pcgym_b200 import make_env, policy_eval  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=65536)
    ap.add_argument("--episodes", type=int, default=5)
    args = ap.parse_args()
    params, _ = bench.cstr_headline(n_envs=args.reps, device="cuda:0")
    params.pop("n_envs")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(3, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                              torch.nn.Linear(64, 1), torch.nn.Tanh()).double().cuda().requires_grad_(False)
    for graph in (False, True):
        ev = policy_eval(make_env, {"mlp": net}, args.reps, params, cuda_graph=graph, as_torch=True)
        ev.get_rollouts()                              # warm-up (+ capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.episodes):
            data = ev.get_rollouts()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.episodes
        steps = args.reps * (params["N"] - 1)
        print(json.dumps({"mode": "cuda_graph" if graph else "eager", "reps": args.reps, "intervals": params["N"] - 1,
                          "ms_per_episode": round(dt * 1e3, 3), "closed_loop_env_steps_per_s": steps / dt,
                          "mean_reward": float(data["mlp"]["r"].mean())}), flush=True)


if __name__ == "__main__":
    main()
