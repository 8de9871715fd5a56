"""In-memory PMVO job: what PMVO.py's __main__ does between "maps loaded" and "Ori3D.mat / Occ3D.mat written"
(PMVO.py:847-872), without the intermediate .npy round trips.  The CLI entry point (PMVO.py at the repo root)
wraps this and writes the reference's files; bench.py times it.

Multi-GPU (one process per GPU, torch.distributed/NCCL): every rank holds all views; the candidate points are
sharded by index for filter_points / forward (no data-path collective, results all-gathered), the kNN queries, the
head filter, the re-scoring and the near-surface medoids are sharded by point (12 B centres are gathered, not the
400 B neighbour lists), and the voxel fusion either runs replicated (its inputs are replicated by then and the whole
fusion costs 0.12 ms, less than any collective) or -- MH_FUSE_DIST=winners, the form a job with rank-private points
needs -- sharded by voxel slab with an all-gather of the per-voxel winners (16 B per occupied voxel; the union is the
all-reduce(SUM) of the disjoint dense volumes, without moving 201 MB of zeros) -- SURVEY.md §8e.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import pmvo as P
from ._lib import MonoHairError, check, lib, ptr, stream_ptr


_PINNED = {}               # reusable pinned result buffers of pmvo_job_host (the caller must consume them before the next call)
_FORCE_SINGLE = False      # tests: run the single-GPU path inside a multi-rank process


def _dist():
    import torch.distributed as dist
    if _FORCE_SINGLE:
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


_SYMM = {}                 # (device index, n) -> (tensor [2, n, 3] in symmetric memory, rendezvous handle)
_SYMM_BROKEN = False       # set once if symmetric memory cannot be set up on this box


def _sweep_mode(dist, dev):
    """"peer" (sweep spread over the ranks through symmetric memory, NVLink stores) or "replicated"."""
    mode = os.environ.get("MH_SWEEP_DIST", "peer")
    if (mode != "peer" or dist is None or dist.get_backend() != "nccl" or torch.device(dev).type != "cuda"
            or dist.get_world_size() > 16):                  # one node: the kernel takes up to 16 peer pointers
        return "replicated"
    return "peer"


def _symm_buffers(n, dev, dist):
    """[2, n, 3] float32 in symmetric memory (row 0 = ori_new, row 1 = center) + the handle carrying every rank's
    pointer; allocated and exchanged once per size (collective: every rank calls it at the same point).  -> None when
    symmetric memory cannot be set up on this box (the ranks agree on that through one all-reduce, once), in which case
    the sweep runs replicated."""
    global _SYMM_BROKEN
    key = (torch.device(dev).index, int(n))
    if key in _SYMM:
        return _SYMM[key]
    if _SYMM_BROKEN:
        return None
    got = None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty((2, int(n), 3), dtype=torch.float32, device=dev)
        got = (t, symm_mem.rendezvous(t, dist.group.WORLD))
    except Exception as e:                                  # noqa: BLE001 -- any failure means "no peer access here"
        import warnings
        warnings.warn(f"symmetric memory unavailable ({type(e).__name__}: {e}); the medoid sweep runs replicated")
    ok = torch.tensor([1 if got is not None else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        _SYMM_BROKEN = True
        return None
    _SYMM.clear()                                           # one size at a time: a new capture replaces the old buffers
    _SYMM[key] = got
    return got


def block_cyclic_index(n, rank, world, block, dev):
    """global indices of the points rank `rank` owns in the distributed sweep, in its processing order."""
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    return idx[(idx // block) % world == rank]


def _shard(n, rank, world):
    """contiguous shard [a,b) of n items for `rank`."""
    per = (n + world - 1) // world
    a = min(rank * per, n)
    return a, min(a + per, n)


def _all_gather_rows(t, n_total, world, dist):
    """all_gather of equally-padded row shards -> first n_total rows."""
    per = (n_total + world - 1) // world
    pad = per - t.size(0)
    if pad > 0:
        t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], 0)
    out = t.new_empty((world * per,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, t.contiguous())
    return out[:n_total]


def filter_stage(pm, cand, n_cov=None):
    """filter_negative_points on device.  cand float32 [N,3] device tensor -> (surface mask, filter mask) over
    the covered prefix (reference chunk arithmetic, §9-R5)."""
    dist = _dist()
    N = cand.size(0)
    if n_cov is None:
        step = 30 if N % 30 == 0 else 31
        n_cov = min(step * (N // 30), N)
    cand = cand[:n_cov]
    if dist is None:
        _, cnt = pm.filter_counters(cand)
    else:
        r, w = dist.get_rank(), dist.get_world_size()
        a, b = _shard(n_cov, r, w)
        _, cnt_local = pm.filter_counters(cand[a:b])
        cnt = _all_gather_rows(cnt_local.t().contiguous(), n_cov, w, dist).t().contiguous()
    return pm.filter_decide(cnt)


def forward_stage(pm, pts):
    dist = _dist()
    if dist is None:
        _, ori, loss, hc = pm.forward(pts)
        return ori, loss, hc
    r, w = dist.get_rank(), dist.get_world_size()
    n = pts.size(0)
    a, b = _shard(n, r, w)
    _, ori, loss, hc = pm.forward(pts[a:b])
    packed = torch.cat([ori, loss[:, None], hc[:, None].float()], 1)
    packed = _all_gather_rows(packed, n, w, dist)
    return packed[:, :3].contiguous(), packed[:, 3].contiguous(), packed[:, 4] > 0.5


def knn_stage(ref, query, k, dev):
    dist = _dist()
    if dist is None:
        return P.knn(ref, query, k, dev)
    r, w = dist.get_rank(), dist.get_world_size()
    n = query.size(0)
    a, b = _shard(n, r, w)
    idx = P.knn(ref, query[a:b].contiguous(), k, dev) if b > a else torch.empty((0, k), dtype=torch.int32, device=dev)
    return _all_gather_rows(idx, n, w, dist)


def head_filter_stage(pm, pts, thr):
    """PMVO.filter_head_points for all points at once (it depends on positions only), sharded over ranks."""
    dist = _dist()
    if dist is None or pts.size(0) == 0:
        return pm.filter_head_points(pts, thr)
    r, w = dist.get_rank(), dist.get_world_size()
    n = pts.size(0)
    a, b = _shard(n, r, w)
    f = pm.filter_head_points(pts[a:b].contiguous(), thr).to(torch.uint8)
    return _all_gather_rows(f, n, w, dist).bool()


def refine_stage(pm, pts, ori, loss, sub_num=5000, k=100):
    """PMVO.refine step (i).  Position-only work is hoisted out of the chunk loop and sharded over ranks: the kNN
    and the head filter.  What the reference makes sequential -- the medoid of the CURRENT neighbour orientations and
    the orientation update, chunk by chunk (Gauss-Seidel across chunks, §9-R7) -- runs as one dependency-ordered sweep
    kernel, replicated on every rank (each needs the final arrays); the re-scoring of (point, medoid) feeds nothing
    back into the sweep, so it runs afterwards over all points at once, sharded over ranks."""
    dev = pm.device
    n = pts.size(0)
    if n == 0:
        return ori.clone(), loss.clone()
    filt = head_filter_stage(pm, pts, pm.visible_threshold).to(torch.uint8).contiguous()
    o_in = ori.contiguous()
    dist = _dist()
    symm = _symm_buffers(n, dev, dist) if _sweep_mode(dist, dev) == "peer" else None
    if symm is not None:
        # the sweep spread over the ranks: each rank queries the neighbours of ITS points only (no 400 B/point
        # all-gather), finished points are stored into every rank's copy over NVLink (mh_refine_sweep_dist)
        r, w = dist.get_rank(), dist.get_world_size()
        mine = block_cyclic_index(n, r, w, int(lib().mh_refine_sweep_dist_block()), dev)
        assert mine.numel() == lib().mh_refine_sweep_dist_local_count(n, r, w)
        nbr = P.knn(pts, pts[mine].contiguous(), k, dev) if mine.numel() else torch.empty((0, k), dtype=torch.int32, device=dev)
        buf, hdl = symm
        buf[0].view(torch.int32).fill_(-1)                  # every word PENDING
        hdl.barrier()                                       # stream-ordered: no peer stores into a copy before its fill
        po = (C.c_uint64 * w)(*[int(b) for b in hdl.buffer_ptrs])
        pc = (C.c_uint64 * w)(*[int(b) + 4 * 3 * n for b in hdl.buffer_ptrs])
        scratch = torch.empty((64,), dtype=torch.uint8, device=dev)
        err = torch.empty((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_refine_sweep_dist(stream_ptr(dev), ptr(o_in), ptr(nbr), k, n, sub_num, r, w, po, pc,
                                             float(os.environ.get("MH_SWEEP_SPIN_S", "2")), 0, ptr(scratch), 64, ptr(err)),
                  "mh_refine_sweep_dist")
        hdl.barrier()                                       # peers store into this copy until their kernels end
        if int(err.item()):
            raise MonoHairError("distributed sweep: a wait on a peer's result ran out (a rank died or never launched)")
        o_new, center = buf[0].clone(), buf[1].clone()
    else:
        nbr = knn_stage(pts, pts, k, dev)
        o_new, center = torch.empty_like(o_in), torch.empty_like(o_in)
        wsb = lib().mh_refine_sweep_workspace_bytes(n, sub_num)
        scratch = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_refine_sweep(stream_ptr(dev), ptr(o_in), ptr(nbr), k, n, sub_num, ptr(o_new), ptr(center),
                                        ptr(scratch), wsb), "mh_refine_sweep")
    if dist is None:
        upd = pm.refine_loss_raw(pts, center)
    else:
        r, w = dist.get_rank(), dist.get_world_size()
        a, b = _shard(n, r, w)
        upd = _all_gather_rows(pm.refine_loss_raw(pts[a:b].contiguous(), center[a:b].contiguous()), n, w, dist).contiguous()
    l = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().mh_refine_finish(stream_ptr(dev), ptr(upd), ptr(filt), n, ptr(l)), "mh_refine_finish")
    return o_new, l


def medoid_stage(ori, nbr_fn, n, dev):
    """medoid orientation of each query's neighbours, sharded by query: nbr_fn(a, b) -> int32 [b-a, K] neighbour
    indices of queries a..b; only the 12 B centres are gathered."""
    dist = _dist()
    if dist is None:
        return P.medoid_gather(ori, nbr_fn(0, n), dev)
    r, w = dist.get_rank(), dist.get_world_size()
    a, b = _shard(n, r, w)
    c = P.medoid_gather(ori, nbr_fn(a, b), dev) if b > a else ori.new_zeros((0, 3))
    return _all_gather_rows(c, n, w, dist).contiguous()


def fuse_stage(pm, pts, dirs, grid=P.GRID, voxel_min=P.VOXEL_MIN, voxel_size=P.VOXEL_SIZE, valid=None, mode=None):
    """Voxel fusion.  Several ranks, mode "replicated" (default: the inputs are replicated at this point of the job):
    every rank fuses everything, no collective.  Mode "winners": each rank fuses the points whose voxel z-slab it owns
    down to the per-voxel winners, the winner lists are all-gathered over NVLink and every rank scatters the union
    into its zero-filled volume."""
    dev = pm.device
    dist = _dist()
    mode = mode or os.environ.get("MH_FUSE_DIST", "replicated")
    if dist is None or mode == "replicated":
        return P.voxel_fuse(pts, dirs, dev, grid, voxel_min, voxel_size, valid=valid)
    r, w = dist.get_rank(), dist.get_world_size()
    gz = int(grid[2])
    # owner by z slab of the voxel index; float64 index math identical to p2v (points[:,2] flipped)
    z = torch.round((-(pts[:, 2].double()) - float(voxel_min[2])) / float(voxel_size)).clamp_(0, gz - 1).long()
    za, zb = _shard(gz, r, w)
    mine = (z >= za) & (z < zb)
    if valid is not None:
        mine &= valid.bool()
    win, cnt = P.voxel_fuse_winners(pts, dirs, dev, grid, voxel_min, voxel_size, valid=mine)
    counts = torch.empty((w,), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(counts, cnt)
    m = max(int(counts.max().item()), 1)                    # one host read: sizes the exchange
    allw = torch.empty((w * m, 4), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(allw, win[:m].contiguous())  # entries past a rank's count carry key -1
    return P.voxel_scatter(allw, dev, grid)


def pmvo_job_device(pm, cand, threshold, stats=None, mark=None, grid=P.GRID, voxel_min=P.VOXEL_MIN, voxel_size=P.VOXEL_SIZE,
                    cand64=None):
    """Whole PMVO job with everything resident on the device.  cand float32 [N,3].  -> dict.
    `mark(name)` (optional) is called at stage boundaries (bench.py records CUDA events there).
    cand64 (optional): the same candidates as loaded, float64 [N,3] -- the reference queries the near-surface neighbours
    with the float64 points and casts them to float32 afterwards (PMVO.py:670-671); without it the float32 ones are used."""
    mark = mark or (lambda name: None)
    mark("start")
    surface, filt = filter_stage(pm, cand)
    mark("filter")
    n_cov = surface.numel()
    pts = cand[:n_cov][surface].contiguous()
    fu = cand[:n_cov][filt].contiguous()
    fu_q = cand64[:n_cov][filt].contiguous() if cand64 is not None else fu
    mark("compact")
    ori, loss, hc = forward_stage(pm, pts)
    mark("optimize")
    o2, l2 = refine_stage(pm, pts, ori, loss)
    mark("refine")
    sel = l2 < threshold
    sp, so = pts[sel].contiguous(), o2[sel].contiguous()
    dev = pm.device
    if fu.size(0) > 0 and sp.size(0) >= 100:
        fh = head_filter_stage(pm, fu, pm.visible_threshold)
        center = medoid_stage(so, lambda a, b: P.knn(sp, fu_q[a:b].contiguous(), 100, dev, cell_factor=0.65), fu.size(0), dev)
        # the head-filtered points are masked out of the fusion instead of being compacted away first: the compaction
        # needs a host synchronisation, which would expose the launch latency of the whole fusion
        all_p, all_o = torch.cat([sp, fu], 0), torch.cat([so, center], 0)
        valid = torch.cat([torch.ones(sp.size(0), dtype=torch.bool, device=dev), ~fh], 0)
    else:
        fh = center = None
        all_p, all_o, valid = sp, so, None
    mark("unvisible")
    vol = fuse_stage(pm, all_p, all_o, grid=grid, voxel_min=voxel_min, voxel_size=voxel_size, valid=valid)
    mark("fuse")
    if center is not None:
        fo, fp = center[~fh], fu[~fh]
    else:
        fo, fp = so.new_zeros((0, 3)), so.new_zeros((0, 3))
    mark("finish")
    out = {"volume": vol, "surface": surface, "filter": filt, "select_p": pts, "select_o": ori, "min_loss": loss,
           "high_conf": hc, "refine_o": o2, "refine_loss": l2, "fu_points": fp, "fu_ori": fo,
           "n_optimized": int(pts.size(0)), "n_selected": int(sp.size(0))}
    if stats is not None:
        stats.update(n_candidates=int(cand.size(0)), n_surface=int(pts.size(0)), n_filter=int(fu.size(0)),
                     n_selected=int(sp.size(0)), n_fu=int(fp.size(0)))
    return out


def pmvo_job_host(camera, depths, Ori, Conf, masks, candidates_host, image_size, patch_size, visible_threshold,
                  conf_threshold, threshold, device="cuda:0", u8=False, readback=None):
    """End to end from HOST buffers to HOST results: H2D of every view's maps (PMVO.__init__), the job, and the
    D2H read of the fused volume and per-point results.  Several ranks: the results are replicated on every GPU and
    rank 0 is the one that hands them to the caller / writes the files, so by default only rank 0 reads them back
    (readback=True forces it everywhere)."""
    if u8:
        pm = P.PMVO.from_u8(camera, depths, Ori, Conf, masks, device=device, image_size=image_size,
                            patch_size=patch_size, visible_threshold=visible_threshold, conf_threshold=conf_threshold)
    else:
        pm = P.PMVO(camera, depths, Ori, Conf, masks, device=device, image_size=image_size, patch_size=patch_size,
                    visible_threshold=visible_threshold, conf_threshold=conf_threshold)
    raw = torch.as_tensor(candidates_host).to(device, non_blocking=True).contiguous()
    cand = raw.type(torch.float).contiguous()
    out = pmvo_job_device(pm, cand, threshold, cand64=raw if raw.dtype == torch.float64 else None)
    host = {}
    if readback is None:
        readback = _dist() is None or _dist().get_rank() == 0
    for k in ("volume", "select_o", "min_loss", "high_conf") if readback else ():
        t = out[k]
        key = (k, tuple(t.shape), t.dtype)
        if key not in _PINNED:
            _PINNED.clear() if len(_PINNED) > 16 else None
            _PINNED[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        _PINNED[key].copy_(t, non_blocking=True)
        host[k] = _PINNED[key]
    torch.cuda.current_stream(torch.device(device)).synchronize()
    host["n_optimized"] = out["n_optimized"]
    return host
