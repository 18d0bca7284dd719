"""CPU: clip sharding + the gloo (world_size 2) version of the cross-rank result check that the
benchmark runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pseldnets_b200 import shard


def test_clip_shard_partitions_in_order():
    for n in (0, 1, 7, 64, 128, 402000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.clip_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        full = torch.randn(n_clips, 7, 11, 64, generator=g)          # stand-in feature maps, same on all ranks
        a, b = shard.clip_shard(n_clips, rank, world)
        table = shard.gather_clip_checksums(full[a:b], n_clips)
        ref = shard.clip_checksums(full)
        ret[rank] = bool(torch.allclose(table, ref, rtol=0, atol=0)) and table.shape == (n_clips, 2)
    finally:
        dist.destroy_process_group()


def test_gather_checksums_gloo_world2():
    for n_clips in (5, 8):
        port = _free_port()
        with mp.Manager() as m:
            ret = m.dict()
            mp.spawn(_worker, args=(2, port, n_clips, ret), nprocs=2, join=True)
            assert ret[0] and ret[1]


def test_single_process_checksums():
    y = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
    t = shard.gather_clip_checksums(y, 2)
    assert np.allclose(t[:, 0].numpy(), [y[0].sum().item(), y[1].sum().item()])
