"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path (monohair_b200/pipeline.py):
contiguous shards, padded all-gather of per-point results, the volume all-reduce with disjoint z-slabs, and the stage
wrappers (forward / filter / head filter / kNN / fusion with a validity mask) against their single-rank results."""
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
    # 4) the stage wrappers themselves, with CPU stand-ins for the kernels (per-point / per-voxel independent functions):
    #    sharded over the two ranks they must return exactly what the single-rank path returns
    from monohair_b200 import pmvo as P

    class FakePM:
        device = torch.device("cpu")
        visible_threshold = 1

        def forward(self, pts):
            o = torch.stack([pts[:, 0] * 2, pts[:, 1] - 1, pts[:, 2] ** 2], 1)
            return pts, o, pts.sum(1), pts[:, 0] > 0

        def filter_counters(self, pts):
            return None, torch.stack([pts[:, 0], pts[:, 1], pts[:, 2], pts.sum(1), pts.prod(1)], 0)

        def filter_decide(self, cnt):
            return cnt[3] > 0, cnt[4] > 0

        def filter_head_points(self, pts, thr):
            return pts[:, 1] > 0.3

        def refine_loss_raw(self, pts, center):
            return (pts * center).sum(1)

    def cpu_knn(ref, query, k, dev):
        d = torch.cdist(query.double(), ref.double())
        return d.topk(k, dim=1, largest=False).indices.int()

    def cpu_fuse(pts, dirs, dev, grid, voxel_min, voxel_size, return_index=False, valid=None):
        gx, gy, gz = [int(v) for v in grid]
        vol = torch.zeros((gz, gy, gx, 4))
        z = torch.round((-(pts[:, 2].double()) - float(voxel_min[2])) / float(voxel_size)).clamp_(0, gz - 1).long()
        keep = torch.ones(pts.size(0), dtype=torch.bool) if valid is None else valid.bool()
        for i in torch.nonzero(keep)[:, 0].tolist():                 # "last point of a slab wins": any per-voxel rule will do
            vol[z[i], 0, 0] = torch.cat([dirs[i], torch.ones(1)])
        return vol

    def cpu_winners(pts, dirs, dev, grid, voxel_min, voxel_size, valid=None, capacity=None):
        gx, gy, gz = [int(v) for v in grid]
        vol = cpu_fuse(pts, dirs, dev, grid, voxel_min, voxel_size, valid=valid).view(-1, 4)
        keys = torch.nonzero(vol[:, 3] > 0)[:, 0]
        cap = pts.size(0)
        win = torch.zeros((cap, 4))
        win[:, 3] = torch.full((cap,), -1, dtype=torch.int32).view(torch.float32)
        win[: keys.numel(), :3] = vol[keys, :3]
        win[: keys.numel(), 3] = keys.int().view(torch.float32)
        return win, torch.tensor([keys.numel()], dtype=torch.int32)

    def cpu_scatter(winners, dev, grid, volume=None):
        gx, gy, gz = [int(v) for v in grid]
        vol = torch.zeros((gz * gy * gx, 4))
        k = winners[:, 3].contiguous().view(torch.int32).long()
        ok = k >= 0
        vol[k[ok], :3] = winners[ok, :3]
        vol[k[ok], 3] = 1.0
        return vol.view(gz, gy, gx, 4)
    P.knn, P.voxel_fuse, P.voxel_fuse_winners, P.voxel_scatter = cpu_knn, cpu_fuse, cpu_winners, cpu_scatter
    P.medoid_gather = lambda ori, nbr, dev: ori[nbr[:, 0].long()].contiguous()
    pm = FakePM()
    g = torch.Generator().manual_seed(5)
    pts = torch.rand((1003, 3), generator=g) - 0.4
    dirs = torch.rand((1003, 3), generator=g)
    valid = torch.rand(1003, generator=g) < 0.7
    grid, vmin, vs = (2, 2, 24), (-0.32, -0.32, -0.6), 0.05
    multi = (PL.forward_stage(pm, pts), PL.filter_stage(pm, pts), PL.head_filter_stage(pm, pts, 1),
             PL.knn_stage(pts[:200], pts, 7, pm.device), PL.fuse_stage(pm, pts, dirs, grid, vmin, vs, valid=valid),
             PL.fuse_stage(pm, pts, dirs, grid, vmin, vs),
             PL.fuse_stage(pm, pts, dirs, grid, vmin, vs, valid=valid, mode="winners"),
             PL.fuse_stage(pm, pts, dirs, grid, vmin, vs, mode="winners"),
             PL.medoid_stage(dirs, lambda a, b: cpu_knn(pts[:200], pts[a:b], 7, pm.device), pts.size(0), pm.device))
    PL._FORCE_SINGLE = True
    single = (PL.forward_stage(pm, pts), PL.filter_stage(pm, pts), PL.head_filter_stage(pm, pts, 1),
              PL.knn_stage(pts[:200], pts, 7, pm.device), PL.fuse_stage(pm, pts, dirs, grid, vmin, vs, valid=valid),
              PL.fuse_stage(pm, pts, dirs, grid, vmin, vs),
              PL.fuse_stage(pm, pts, dirs, grid, vmin, vs, valid=valid), PL.fuse_stage(pm, pts, dirs, grid, vmin, vs),
              PL.medoid_stage(dirs, lambda a, b: cpu_knn(pts[:200], pts[a:b], 7, pm.device), pts.size(0), pm.device))
    PL._FORCE_SINGLE = False

    def same(a, b):
        if isinstance(a, (tuple, list)):
            return all(same(x, y) for x, y in zip(a, b))
        return torch.equal(a, b)
    for name, m, s1 in zip(("forward", "filter", "head_filter", "knn", "fuse(valid)", "fuse", "fuse winners(valid)", "fuse winners",
                            "medoid"), multi, single):
        assert same(m, s1), f"{name}_stage: sharded result differs from the single-rank result"
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
