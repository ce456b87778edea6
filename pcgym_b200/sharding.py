"""Multi-GPU plumbing: the env batch shards embarrassingly, one process per GPU (torchrun).

There is NO collective on the step path: every rank owns a contiguous env range and its own
descriptor (same seed, `env_id0` = first global env id so the Philox noise streams of a sharded run
are identical to the single-GPU run).  Collectives appear only (a) in the benchmark's barrier /
max-over-ranks timing and (b) in the optional `all_gather_envs`, which concatenates per-rank rollout
buffers along the env axis over NCCL (NVLink 5 / NVSwitch) when one consumer needs them all.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced split; every shard is even-sized when possible (128-bit row alignment)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_params(env_params: dict, rank: int, world: int) -> dict:
    """env_params for this rank's shard of a global batch `n_envs`: slices per-env tensors, sets env_id0."""
    p = dict(env_params)
    n = int(p.get("n_envs", 1))
    lo, hi = shard_range(n, rank, world)
    p["n_envs"], p["env_id0"] = hi - lo, int(p.get("env_id0", 0)) + lo

    def cut(v):
        return v[lo:hi] if hasattr(v, "ndim") and v.ndim == 2 and v.shape[0] == n else v

    if "x0" in p:
        p["x0"] = cut(p["x0"])
    for key in ("SP", "disturbances"):
        if p.get(key) is not None:
            p[key] = {k: cut(v) for k, v in p[key].items()}
    return p


def mixed_plan(counts: dict, rank: int, world: int) -> dict:
    """Config-5 style mixed batches: split EACH model's envs evenly so every GPU carries the same mix."""
    return {name: shard_range(n, rank, world) for name, n in counts.items()}


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":                      # bind the communicator to this rank's GPU (barrier() then needs no guess)
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int, local_world: int = 1) -> list:
    """Pin this process to CPUs close to GPU `local_rank` (NVML's affinity mask = its NUMA node), so that pinned host
    buffers allocated afterwards are first-touched next to the GPU's PCIe root.  When several local GPUs report the
    SAME mask (one socket serving all of them) the mask is cut into disjoint, contiguous slices, one per rank, so the
    ranks' Python launch threads never time-share a core (round 1 pinned all 8 ranks onto the same 32 CPUs).
    Returns the CPU list it bound to ([] = left unchanged: NVML unavailable or mask empty)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        n_words = (os.cpu_count() + 63) // 64
        allowed = os.sched_getaffinity(0)

        def mask_of(g):
            m = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(g)), n_words)
            return tuple(c for c in (64 * w + b for w, word in enumerate(m) for b in range(64) if (int(word) >> b) & 1)
                         if c in allowed)

        mine = mask_of(local_rank)
        if not mine:
            return []
        peers = [g for g in range(max(int(local_world), 1)) if mask_of(g) == mine]
        cpus = slice_cpus(list(mine), peers.index(local_rank) if local_rank in peers else 0, max(len(peers), 1))
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def slice_cpus(cpus: list, i: int, m: int) -> list:
    """Slice i of m of a CPU list: contiguous, disjoint, at least one CPU each (falls back to sharing when m > len)."""
    if m <= 1 or len(cpus) < m:
        return list(cpus)
    lo, hi = (i * len(cpus)) // m, ((i + 1) * len(cpus)) // m
    return list(cpus[lo:hi])


def gather_over_ranks(values, device=None) -> list:
    """[[v0, v1, ...] of rank 0, of rank 1, ...]: small per-rank diagnostics (kernel time, enqueue time) for rank 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [[float(v) for v in values]]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64,
                     device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [[float(v) for v in o.tolist()] for o in outs]


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reduced with MAX so the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_gather_envs(local: torch.Tensor, sizes=None) -> torch.Tensor:
    """Concatenate SoA rollout buffers [..., n_local] from every rank along the env (last) axis.
    Off the timed path: a K=99, 65 536-env CSTR rollout is ~0.5 GB per GPU."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if sizes is None:
        n = torch.tensor([local.shape[-1]], dtype=torch.int64, device=local.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n)
        sizes = [int(v.item()) for v in all_n]
    env_major = local.movedim(-1, 0).contiguous()                      # [n_local, ...]
    n_max = max(sizes)
    if env_major.shape[0] < n_max:                                     # shards differ by at most one env: pad
        pad = torch.zeros((n_max - env_major.shape[0],) + tuple(env_major.shape[1:]), dtype=local.dtype,
                          device=local.device)
        env_major = torch.cat([env_major, pad], dim=0)
    outs = [torch.empty_like(env_major) for _ in sizes]
    dist.all_gather(outs, env_major)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0).movedim(0, -1)
