"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path (monohair_b200/pipeline.py):
contiguous shards, padded all-gather of per-point results, and the volume all-reduce with disjoint z-slabs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monohair_b200 import pipeline as PL
    assert PL._dist() is dist
    # 1) shards tile [0,n) exactly, in rank order
    n = 1003
    a, b = PL._shard(n, rank, world)
    cover = torch.zeros(n)
    cover[a:b] = 1
    dist.all_reduce(cover)
    assert torch.all(cover == 1)
    # 2) padded all-gather restores the global order (per-point forward results)
    full = torch.arange(n * 5, dtype=torch.float32).reshape(n, 5)
    got = PL._all_gather_rows(full[a:b].clone(), n, world, dist)
    assert torch.equal(got, full)
    idx = torch.arange(n * 3, dtype=torch.int32).reshape(n, 3)
    assert torch.equal(PL._all_gather_rows(idx[a:b].clone(), n, world, dist), idx)
    # 3) volume fusion: z-slab ownership is a partition, and SUM of the per-rank volumes equals the union exactly
    gz = 192
    za, zb = PL._shard(gz, rank, world)
    rng = np.random.default_rng(0)
    vol_full = torch.from_numpy(rng.normal(size=(gz, 4, 4, 4)).astype(np.float32))
    mine = torch.zeros_like(vol_full)
    mine[za:zb] = vol_full[za:zb]
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    assert torch.equal(mine, vol_full)              # disjoint support: bit-exact
    q.put((rank, "ok"))
    dist.destroy_process_group()


def test_world2_gloo_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == [0, 1]
