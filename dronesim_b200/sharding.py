"""Multi-GPU layout: environments are independent, so the swarm shards by contiguous env ranges
with NO per-step communication.  The only collective is the end-of-rollout statistics all-reduce.

The reference is single-process (its only "parallelism" is ``for i in range(self.NUM_DRONES)``,
dronesim/envs/BaseAviary.py:522); intra-env coupling (downwash BaseAviary.py:1747-1763, adjacency
:913-921) never crosses an env, hence never crosses a GPU.

Integer maps here are exact and identical for any world size:
    global env  e  ->  (rank, local env) = (e // envs_per_rank, e % envs_per_rank)        (even split)
    global vehicle v = e * D + slot
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

SUM_KEYS = ("control_evals", "sum_pos_err_sq", "saturated_cmds", "wls_slow_path", "wls_non_converged", "non_finite",
            "done_vehicles")
MIN_KEYS = ("min_altitude",)


def shard_envs(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(env_offset, n_envs) of ``rank``.  Envs are dealt in contiguous blocks; when the split is
    uneven the first ``total_envs % world_size`` ranks hold one extra env."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_envs, world_size)
    n = base + (1 if rank < extra else 0)
    off = rank * base + min(rank, extra)
    return off, n


def owner_of_env(env: np.ndarray, total_envs: int, world_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised inverse of ``shard_envs``: global env -> (rank, local env)."""
    env = np.asarray(env, dtype=np.int64)
    base, extra = divmod(total_envs, world_size)
    cut = extra * (base + 1)
    rank = np.where(env < cut, env // max(base + 1, 1), extra + (env - cut) // max(base, 1))
    off = rank * base + np.minimum(rank, extra)
    return rank.astype(np.int64), (env - off).astype(np.int64)


def vehicle_to_env_slot(v: np.ndarray, drones_per_env: int) -> Tuple[np.ndarray, np.ndarray]:
    v = np.asarray(v, dtype=np.int64)
    return v // drones_per_env, v % drones_per_env


def allreduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """One tiny all-reduce (sum) + one (min) over the rollout statistics.  NCCL on GPUs, gloo on CPU."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    s = torch.tensor([float(stats[k]) for k in SUM_KEYS], dtype=torch.float64, device=dev)
    m = torch.tensor([float(stats[k]) for k in MIN_KEYS], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MIN, group=group)
    out = {k: float(s[i]) for i, k in enumerate(SUM_KEYS)}
    out.update({k: float(m[i]) for i, k in enumerate(MIN_KEYS)})
    return out
