"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: env sharding + reward all-gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pde_opt_b200.parallel import gather_env_values, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_envs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(num_envs, rank, world)
    local = torch.stack([torch.arange(lo, hi, dtype=torch.float32), torch.arange(lo, hi, dtype=torch.float32) ** 2], 1)
    full = gather_env_values(local, num_envs)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_envs", [8, 7])
def test_env_sharding_and_reward_gather(num_envs):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([np.arange(num_envs, dtype=np.float32), np.arange(num_envs, dtype=np.float32) ** 2], 1)
    for r in range(world):
        np.testing.assert_array_equal(res[r], expect)


def test_shard_bounds_cover_and_balance():
    for n in [1, 7, 8, 4096, 4097]:
        for w in [1, 2, 4, 8]:
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
