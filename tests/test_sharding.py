"""Multi-GPU plumbing on CPU: shard arithmetic, and world_size-2 gloo runs of the timing reduction and
the optional env-axis all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcgym_b200 import sharding


def test_shard_range_covers_and_balances():
    for n, w in [(65536, 8), (524288, 8), (10, 4), (7, 8), (1, 1)]:
        cuts = [sharding.shard_range(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 3, 2)


def test_shard_params_slices_per_env_inputs():
    n = 10
    p = {"n_envs": n, "x0": np.arange(n * 3.0).reshape(n, 3), "SP": {"Ca": np.ones((n, 5)), "T": [1.0] * 5},
         "disturbances": {"Ti": np.arange(n * 5.0).reshape(n, 5)}, "seed": 3}
    s = sharding.shard_params(p, 1, 4)
    lo, hi = sharding.shard_range(n, 1, 4)
    assert s["n_envs"] == hi - lo and s["env_id0"] == lo
    np.testing.assert_array_equal(s["x0"], p["x0"][lo:hi])
    assert s["SP"]["Ca"].shape == (hi - lo, 5) and s["SP"]["T"] == [1.0] * 5
    np.testing.assert_array_equal(s["disturbances"]["Ti"], p["disturbances"]["Ti"][lo:hi])
    assert p["n_envs"] == n and "env_id0" not in p


def test_mixed_plan_keeps_the_mix_per_gpu():
    counts = {"cstr": 174763, "four_tank": 174763, "crystallization": 174762}
    per = [sharding.mixed_plan(counts, r, 8) for r in range(8)]
    for name, n in counts.items():
        assert sum(b - a for a, b in (pl[name] for pl in per)) == n
        sizes = [pl[name][1] - pl[name][0] for pl in per]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = sharding.init_from_env("gloo")
    assert (r, w) == (rank, world)
    ms = sharding.max_over_ranks(10.0 + 5.0 * rank, "cpu")
    total = sharding.sum_over_ranks(float(rank + 1), "cpu")
    lo, hi = sharding.shard_range(11, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64).repeat(3, 2, 1)          # [K=3, field=2, n_local]
    full = sharding.all_gather_envs(local)
    q.put((rank, ms, total, full.shape, full[0, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_timing_and_gather():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, total, shape, row in res:
        assert ms == 15.0 and total == 3.0                 # MAX over ranks for time, SUM for work
        assert tuple(shape) == (3, 2, 11) and row == list(map(float, range(11)))


def test_single_process_passthrough():
    assert sharding.max_over_ranks(3.5) == 3.5
    t = torch.ones(2, 3)
    assert sharding.all_gather_envs(t) is t
