"""CPU: multi-GPU bookkeeping - integer maps are exact and world-size invariant; the one collective of
the path (end-of-rollout stats all-reduce) is exercised with world_size = 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from dronesim_b200 import sharding as S
from dronesim_b200.workloads import hetero16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("total,world", [(16, 1), (16, 2), (17, 4), (65536, 8), (3, 8), (0, 2)])
def test_shard_envs_partition(total, world):
    seen = []
    for r in range(world):
        off, n = S.shard_envs(total, world, r)
        seen.extend(range(off, off + n))
    assert seen == list(range(total))
    if total:
        env = np.arange(total)
        rank, local = S.owner_of_env(env, total, world)
        for e in range(total):
            off, n = S.shard_envs(total, world, int(rank[e]))
            assert off + local[e] == e and 0 <= local[e] < n


def test_vehicle_env_slot_map():
    v = np.arange(0, 64)
    e, s = S.vehicle_to_env_slot(v, 16)
    np.testing.assert_array_equal(e * 16 + s, v)
    assert e.dtype == np.int64 and s.max() == 15


def test_workload_is_sharding_invariant():
    """The synthetic swarm is a pure function of the global env index: any split reproduces it bit for bit."""
    _, _, _, pos_all, act_all, tgt_all = hetero16(24, seed=0)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            off, n = S.shard_envs(24, world, r)
            parts.append(hetero16(n, seed=0, env_offset=off)[3])
        np.testing.assert_array_equal(np.concatenate(parts, axis=0), pos_all)
    z = pos_all[0, :, 2]
    assert (np.diff(np.sort(z)) > 0.2).all()  # distinct altitudes: the downwash model is singular at dz -> 0+


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from dronesim_b200 import sharding as S2

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {k: float(rank + 1) * (i + 1) for i, k in enumerate(S2.SUM_KEYS)}
    stats["min_altitude"] = 0.5 - 0.1 * rank
    out = S2.allreduce_stats(stats)
    off, n = S2.shard_envs(10, world, rank)
    q.put((rank, out, off, n))
    dist.destroy_process_group()


def test_allreduce_stats_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for rank, out, off, n in res:
        for i, k in enumerate(S.SUM_KEYS):
            assert out[k] == 3.0 * (i + 1)
        assert out["min_altitude"] == 0.4
    assert [(r[2], r[3]) for r in res] == [(0, 5), (5, 5)]
