"""Host-side multi-rank logic on CPU: site sharding and the gloo all-reduce of the
statistic buffers (world_size 2)."""
import os
import socket

import numpy as np
import pytest

from nxblink_b200 import distributed as nd


def test_shard_range_partitions():
    for n in (0, 1, 7, 393, 15625):
        for world in (1, 2, 3, 8):
            blocks = [nd.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    counts = rng.integers(0, 1 << 40, size=1000).astype(np.uint64)
    dwell = rng.random(500)
    c, d = nd.allreduce_host_summary(counts, dwell)
    if rank == 0:
        np.save(out + '.c.npy', c)
        np.save(out + '.d.npy', d)
    dist.destroy_process_group()


def test_gloo_allreduce_of_summary_buffers(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    out = str(tmp_path / 'res')
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    exp_c = sum(np.random.default_rng(100 + r).integers(0, 1 << 40, size=1000).astype(np.uint64)
            for r in range(world))
    exp_d = np.zeros(500)
    for r in range(world):
        g = np.random.default_rng(100 + r)
        g.integers(0, 1 << 40, size=1000)
        exp_d += g.random(500)
    assert (np.load(out + '.c.npy') == exp_c).all()         # integer statistics: exact
    np.testing.assert_allclose(np.load(out + '.d.npy'), exp_d, rtol=1e-15)
