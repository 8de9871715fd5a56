"""B200-native PMVO: host-side mirror of the reference's PMVO.py (same class / function names, argument meaning,
file outputs) driving the CUDA kernels of libmonohair_b200.so through the C ABI.  No CPU fallback.

Reference: /root/reference/PMVO.py  (class PMVO :13-529, filter_negative_points :535-557, optimize :565-595,
refine :602-764, config_parser :767-800).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import MhViews, check, lib, ptr, stream_ptr

# module globals the reference reads inside its functions: set them like PMVO.py:809-820 does
bust_tree = None
scalp_tree = None          # anything with a `.data` [M,3] float64 array (e.g. scipy KDTree) or an ndarray
scalp_max = None
device = "cuda:0"
Num_points = None
args = None

VOXEL_MIN = np.array([-0.32, -0.32, -0.24])     # PMVO.py:699
VOXEL_SIZE = 0.005 / 2                          # PMVO.py:700
GRID = (256, 256, 192)                          # PMVO.py:695


def _f32(a, dev, non_blocking=False):
    t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
    return t.to(dev, non_blocking=non_blocking).type(torch.float).contiguous()


def _world():
    """(dist module, rank, world) when torch.distributed is initialised with more than one rank, else (None, 0, 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and SHARD_VIEW_UPLOAD:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _rank0():
    """True on the rank that writes the job's files (every rank when not distributed)."""
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def _barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def broadcast_array(a, dev, dtype=np.float64):
    """rank 0's numpy array on every rank (shape first, then the data through the device)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return a
    shp = torch.zeros((4,), dtype=torch.int64, device=dev)
    if dist.get_rank() == 0:
        a = np.ascontiguousarray(a, dtype=dtype)
        shp[0] = a.ndim
        shp[1:1 + a.ndim] = torch.tensor(a.shape, dtype=torch.int64)
    dist.broadcast(shp, 0)
    nd = int(shp[0].item())
    shape = [int(v) for v in shp[1:1 + nd].tolist()]
    t = torch.from_numpy(a).to(dev) if dist.get_rank() == 0 else torch.empty(shape, dtype=torch.from_numpy(np.zeros(1, dtype)).dtype, device=dev)
    dist.broadcast(t, 0)
    return t.cpu().numpy()


SHARD_VIEW_UPLOAD = True   # multi-GPU: each rank uploads + packs only its block of views, planes are all-gathered over NVLink


def _alloc_planes(V, H, W, dev):
    """-> (mapC, mapP, first view of this rank, one-past-last view, finish())."""
    dist, r, w = _world()
    if dist is None:
        mapC = torch.empty((V, H, W, 4), dtype=torch.float32, device=dev)
        mapP = torch.empty((V, H, W, 4), dtype=torch.float32, device=dev)
        return mapC, mapP, 0, V, 0, (lambda: (mapC, mapP))
    per = (V + w - 1) // w
    a, b = min(r * per, V), min((r + 1) * per, V)
    locC = torch.zeros((per, H, W, 4), dtype=torch.float32, device=dev)
    locP = torch.zeros((per, H, W, 4), dtype=torch.float32, device=dev)

    def finish():
        fullC = torch.empty((w * per, H, W, 4), dtype=torch.float32, device=dev)
        fullP = torch.empty((w * per, H, W, 4), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(fullC, locC)
        dist.all_gather_into_tensor(fullP, locP)
        return fullC[:V], fullP[:V]
    return locC, locP, a, b, a, finish


class PMVO(nn.Module):
    """Drop-in for the reference class (PMVO.py:13-37).  The per-view maps are packed once into two resident
    planes per view (see csrc/pmvo_views.cu); everything else runs in fused kernels."""

    def __init__(self, camera, depths, Ori, Conf, masks, device='cuda:0', image_size=[1120, 1992], patch_size=5,
                 visible_threshold=1, conf_threshold=0.4):
        super().__init__()
        lib()
        if not torch.cuda.is_available():
            raise _lib.MonoHairError("monohair_b200.PMVO needs a CUDA device (there is no CPU fallback)")
        self.camera_dict = camera
        self.visible_threshold = visible_threshold
        self.camera = list(camera.values())
        self.camera_key = list(camera.keys())
        self.device = torch.device(device)
        self.image_size = list(image_size)
        self.patch_size = int(patch_size)
        self.conf_threshold = conf_threshold
        H, W = int(image_size[0]), int(image_size[1])
        V = len(self.camera)
        self.V, self.H, self.W = V, H, W
        dev = self.device
        with torch.cuda.device(dev):
            mapC, mapP, v_lo, v_hi, v_off, finish = _alloc_planes(V, H, W, dev)
            self.cam = torch.stack([c.record() for c in self.camera]).to(dev).contiguous()
            st = stream_ptr(dev)
            stage = {}

            def to_dev(a):
                """raw H2D copy into a reusable staging buffer (async when the source is pinned): the copy of
                PMVO.py:23-26; the float cast is fused into the pack kernel."""
                t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
                key = (t.dtype, tuple(t.shape))
                if key not in stage:
                    stage[key] = [torch.empty(t.shape, dtype=t.dtype, device=dev) for _ in range(2)]
                buf = stage[key][0]
                stage[key].reverse()
                buf.copy_(t, non_blocking=True)
                return buf

            for v, k in enumerate(self.camera_key):
                if not (v_lo <= v < v_hi):
                    continue                                  # another rank uploads this view
                d_h, o_h, c_h, m_h = depths[k], Ori[k], Conf[k], masks[k]
                dts = [(x.dtype if torch.is_tensor(x) else torch.from_numpy(np.asarray(x)[:0].copy()).dtype) for x in (d_h, o_h, c_h, m_h)]
                f64_path = dts[0] == torch.float32 and all(t == torch.float64 for t in dts[1:])
                if f64_path:
                    d, o, c, m = to_dev(d_h), to_dev(o_h), to_dev(c_h), to_dev(m_h)
                else:
                    d, o, c, m = _f32(d_h, dev), _f32(o_h, dev), _f32(c_h, dev), _f32(m_h, dev)
                assert d.shape[:2] == (H, W) and o.shape == (H, W, 2) and c.shape == (H, W) and m.shape[:2] == (H, W), \
                    f"view {k}: map shapes do not match image_size {image_size}"
                ds = d.shape[2] if d.dim() == 3 else 1
                ms = m.shape[2] if m.dim() == 3 else 1
                fn = lib().mh_views_pack_f64 if f64_path else lib().mh_views_pack
                check(fn(st, v - v_off, H, W, self.patch_size, ptr(d), ds, ptr(o), ptr(c), ptr(m), ms,
                         ptr(mapC), ptr(mapP)), "mh_views_pack")
            self.mapC, self.mapP = finish()
        self._views = MhViews(V, H, W, self.patch_size, self.mapC.data_ptr(), self.mapP.data_ptr(), self.cam.data_ptr())
        self._offsets = self._sample_offsets(90).to(dev)

    # ------------------------------------------------------------------ construction from the on-disk formats
    @classmethod
    def from_u8(cls, camera, depth, ori_gray, conf_u8, mask_u8, device='cuda:0', image_size=[1120, 1992],
                patch_size=5, visible_threshold=1, conf_threshold=0.4):
        """Fast path: maps in their file formats -- depth float32 [V,H,W] (channel 0 of
        render_depth/*.npy), best_ori gray uint8, conf uint8, hair_mask uint8 (one channel) -- decoded inside
        the pack kernel with lookup tables built by the same numpy expressions as Load_Ori_And_Conf /
        load_mask (PMVO_utils.py:265-272, 302-304).  Inputs may be pinned host tensors / ndarrays."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        lib()
        self.camera_dict = camera
        self.visible_threshold = visible_threshold
        self.camera = list(camera.values())
        self.camera_key = list(camera.keys())
        self.device = torch.device(device)
        self.image_size = list(image_size)
        self.patch_size = int(patch_size)
        self.conf_threshold = conf_threshold
        H, W = int(image_size[0]), int(image_size[1])
        V = len(self.camera)
        self.V, self.H, self.W = V, H, W
        dev = self.device
        g = np.arange(256, dtype=np.uint8)
        o = (180 - g) / 180 * math.pi
        ori_lut = torch.from_numpy(np.stack([np.sin(o), np.cos(o)], -1)).type(torch.float).to(dev).contiguous()
        conf_lut = torch.from_numpy(g / 255.).type(torch.float).to(dev)
        mm = g.copy()
        mm[mm < 50] = 0
        mask_lut = torch.from_numpy(mm / 255.).type(torch.float).to(dev)
        with torch.cuda.device(dev):
            mapC, mapP, v_lo, v_hi, v_off, finish = _alloc_planes(V, H, W, dev)
            self.cam = torch.stack([c.record() for c in self.camera]).to(dev).contiguous()
            st = stream_ptr(dev)
            as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            d_all, o_all, c_all, m_all = as_t(depth), as_t(ori_gray), as_t(conf_u8), as_t(mask_u8)
            assert d_all.dtype == torch.float32 and o_all.dtype == torch.uint8 and c_all.dtype == torch.uint8 \
                and m_all.dtype == torch.uint8
            assert d_all.shape == (V, H, W) and o_all.shape == (V, H, W) and c_all.shape == (V, H, W) and m_all.shape == (V, H, W)
            d_all = d_all[v_lo:v_hi].to(dev, non_blocking=True)
            o_all = o_all[v_lo:v_hi].to(dev, non_blocking=True)
            c_all = c_all[v_lo:v_hi].to(dev, non_blocking=True)
            m_all = m_all[v_lo:v_hi].to(dev, non_blocking=True)
            for v in range(v_lo, v_hi):
                j = v - v_lo
                check(lib().mh_views_pack_u8(st, v - v_off, H, W, self.patch_size, ptr(d_all[j]), 1, ptr(o_all[j]),
                                             ptr(c_all[j]), ptr(m_all[j]), ptr(ori_lut), ptr(conf_lut), ptr(mask_lut),
                                             ptr(mapC), ptr(mapP)), "mh_views_pack_u8")
            self.mapC, self.mapP = finish()
            self._keep = (ori_lut, conf_lut, mask_lut)
        self._views = MhViews(V, H, W, self.patch_size, self.mapC.data_ptr(), self.mapP.data_ptr(), self.cam.data_ptr())
        self._offsets = self._sample_offsets(90).to(dev)
        return self

    @staticmethod
    def _sample_offsets(num_sample=90):
        """The depth offsets of sample_next_3d_pos, built with the reference's torch.arange calls (PMVO.py:274-278)."""
        s1 = torch.arange(-0.005, -0.001, 0.004 / (num_sample / 4))
        s2 = torch.arange(-0.001, 0.001, 0.002 / (num_sample / 2))
        s3 = torch.arange(0.001, 0.005, 0.004 / (num_sample / 4))
        return torch.cat([s1, s2, s3], 0)[:num_sample].contiguous()

    def _vp(self):
        return C.byref(self._views)

    def _pts(self, points):
        if isinstance(points, np.ndarray):
            points = torch.from_numpy(points)
        return points.to(self.device).type(torch.float).contiguous()

    # ------------------------------------------------------------------ PMVO.forward (PMVO.py:39-78)
    def forward(self, points, debug=False):
        pts = self._pts(points)
        N = pts.size(0)
        dev = self.device
        ori = torch.empty((N, 3), dtype=torch.float32, device=dev)
        loss = torch.empty((N,), dtype=torch.float32, device=dev)
        hc = torch.empty((N,), dtype=torch.uint8, device=dev)
        ws = torch.empty((int(lib().mh_pmvo_optimize_workspace_bytes(self._vp(), N)),), dtype=torch.uint8, device=dev)
        dbg = {}
        if debug:
            dbg = {"base_idx": torch.empty((_lib.MH_TOPK, N), dtype=torch.int32, device=dev),
                   "base_val": torch.empty((_lib.MH_TOPK, N), dtype=torch.float32, device=dev),
                   "best_sample": torch.empty((N, 3), dtype=torch.float32, device=dev),
                   "loss_b": torch.empty((_lib.MH_NUM_BASE, N), dtype=torch.float32, device=dev),
                   "arg_b": torch.empty((_lib.MH_NUM_BASE, N), dtype=torch.int32, device=dev)}
        with torch.cuda.device(dev):
            check(lib().mh_pmvo_optimize(stream_ptr(dev), self._vp(), ptr(pts), N, ptr(self._offsets),
                                         self._offsets.numel(), float(self.conf_threshold), ptr(ori), ptr(loss), ptr(hc),
                                         ptr(dbg.get("base_idx")), ptr(dbg.get("base_val")), ptr(dbg.get("best_sample")),
                                         ptr(dbg.get("loss_b")), ptr(dbg.get("arg_b")), ptr(ws), ws.numel()),
                  "mh_pmvo_optimize")
        if debug:
            return pts, ori, loss, hc.bool(), dbg
        return pts, ori, loss, hc.bool()

    # ------------------------------------------------------------------ PMVO.filter_points (PMVO.py:402-459)
    def filter_counters(self, points):
        pts = self._pts(points)
        N = pts.size(0)
        cnt = torch.empty((5, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().mh_filter_count(stream_ptr(self.device), self._vp(), ptr(pts), N, float(self.visible_threshold),
                                        float(self.conf_threshold), ptr(cnt)), "mh_filter_count")
        return pts, cnt

    def filter_decide(self, cnt):
        N = cnt.size(1)
        surface = torch.empty((N,), dtype=torch.uint8, device=self.device)
        filt = torch.empty((N,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().mh_filter_decide(stream_ptr(self.device), ptr(cnt), N, ptr(surface), ptr(filt)), "mh_filter_decide")
        return surface.bool(), filt.bool()

    def filter_points(self, points):
        pts, cnt = self.filter_counters(points)
        surface, filt = self.filter_decide(cnt)
        return surface, pts[surface], filt

    # ------------------------------------------------------------------ PMVO.compute_unvisible_points (:461-480)
    def compute_unvisible_points(self, points):
        pts = self._pts(points)
        N = pts.size(0)
        cnt = torch.empty((N,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().mh_visible_count(stream_ptr(self.device), self._vp(), ptr(pts), N, 0.9, ptr(cnt)), "mh_visible_count")
        return ~(cnt > 2)

    # ------------------------------------------------------------------ PMVO.filter_head_points (:96-144)
    def filter_head_points(self, points, visible_threshold):
        pts = self._pts(points)
        N = pts.size(0)
        dev = self.device
        sc = scalp_tree if isinstance(scalp_tree, np.ndarray) or scalp_tree is None else scalp_tree.data
        if sc is None or scalp_max is None:
            raise _lib.MonoHairError("filter_head_points needs the module globals scalp_tree / scalp_max (PMVO.py:99-106)")
        if not hasattr(self, "_scalp") or self._scalp_src is not sc:
            self._scalp = torch.from_numpy(np.ascontiguousarray(np.asarray(sc, dtype=np.float64))).to(dev)
            self._scalp_src = sc
        cnt = torch.empty((2, N), dtype=torch.float32, device=dev)
        dist = torch.empty((N,), dtype=torch.float64, device=dev)
        filt = torch.empty((N,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib().mh_nn_dist(st, ptr(self._scalp), self._scalp.size(0), ptr(pts), N, ptr(dist)), "mh_nn_dist")
            check(lib().mh_head_count(st, self._vp(), ptr(pts), N, float(visible_threshold), ptr(cnt)), "mh_head_count")
            check(lib().mh_head_decide(st, ptr(cnt), ptr(dist), ptr(pts), N, 0.04, float(scalp_max[2]) - 0.01, ptr(filt)),
                  "mh_head_decide")
        return filt.bool()

    # ------------------------------------------------------------------ PMVO.refine (:81-93)
    def refine_loss_raw(self, points, ori):
        pts = self._pts(points)
        o = ori.to(self.device).type(torch.float).contiguous()
        N = pts.size(0)
        loss = torch.empty((N,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().mh_pmvo_refine_loss(stream_ptr(self.device), self._vp(), ptr(pts), ptr(o), N,
                                            float(self.conf_threshold), ptr(loss)), "mh_pmvo_refine_loss")
        return loss

    def refine(self, points, ori):
        filt = self.filter_head_points(points, self.visible_threshold)
        loss = self.refine_loss_raw(points, ori)
        loss[filt] = -1
        return loss

    # ------------------------------------------------------------------ PMVO.Compute_Visible_and_Ori (:346-376)
    def Compute_Visible_and_Ori(self, points):
        """Sets visible [V,N], Ori [V,N,2], Conf [V,N], mask [V,N].  The patch tensors Ori_patch / Conf_patch of
        the reference are never materialised here: the kernels gather patches straight from the resident maps."""
        pts = self._pts(points)
        N = pts.size(0)
        dev = self.device
        self.visible = torch.empty((self.V, N), dtype=torch.float32, device=dev)
        self.Ori = torch.empty((self.V, N, 2), dtype=torch.float32, device=dev)
        self.Conf = torch.empty((self.V, N), dtype=torch.float32, device=dev)
        self.mask = torch.empty((self.V, N), dtype=torch.float32, device=dev)
        self._rowcol = torch.empty((self.V, N, 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_centre_gather(stream_ptr(dev), self._vp(), ptr(pts), N, ptr(self.visible), ptr(self.Ori),
                                         ptr(self.Conf), ptr(self.mask), ptr(self._rowcol)), "mh_centre_gather")

    def project_points(self, points, camera, image_size=None):
        """PMVO.project_points (:378-397) for one camera: (uv[row,col] int64, z, unvisible_index)."""
        v = self.camera.index(camera) if not isinstance(camera, int) else camera
        self.Compute_Visible_and_Ori(points)
        rc = self._rowcol[v].long()
        oob = rc[:, 1] < 0
        col = torch.where(oob, -rc[:, 1] - 1, rc[:, 1])
        # z = -z_cam / 2 (PMVO.py:396) with z_cam the third row of pose @ [p;1]: the fp32 FMA chain of the kernels
        # (mh_world_to_cam), evaluated here through float64 (each product is exact there), then rounded per step
        pts = self._pts(points).double()
        P = self.cam[v, 8:12].double().cpu().numpy()
        z = torch.zeros_like(pts[:, 0])
        for k in range(3):
            z = (z + P[k] * pts[:, k]).float().double()
        z = (z + P[3]).float()
        return torch.stack([rc[:, 0], col], 1), z / -2.0, oob


# ====================================================================================== module functions
def filter_negative_points(points, pmvo, args, step=30):
    """PMVO.py:535-557, including its chunk arithmetic: N//30 points per chunk and 30 (or 31)
    chunks, so a tail of N - 31*(N//30) points can be dropped exactly as in the reference.  Chunks are
    independent, so the covered prefix is processed in one launch."""
    if points.shape[0] % step != 0:
        step = step + 1
    num_sub_p = points.shape[0] // 30
    n_cov = min(step * num_sub_p, points.shape[0])
    from . import pipeline
    cand = torch.from_numpy(points[:n_cov]).to(args.device).type(torch.float)
    surface_index, filter_index = pipeline.filter_stage(pmvo, cand, n_cov=n_cov)      # point-sharded under torchrun
    surface_points = cand[surface_index]
    surface_points = surface_points.cpu().numpy()
    filter_indexs = filter_index.cpu().numpy()
    surface_indexs = surface_index.cpu().numpy()
    print('surface_num:', surface_points.shape[:])
    print('num filter_unvisible:', np.sum(filter_indexs))
    return surface_indexs, surface_points, filter_indexs


def optimize(points, pmvo, args, chunk=1 << 20):
    """PMVO.py:565-595: forward over all points (the 5000-point chunking of the reference only bounds its
    [V,N,90] temporaries; points are independent) and the four .npy dumps."""
    outs = [[], [], [], []]
    n_total = Num_points if Num_points is not None else points.shape[0]
    for i in range(0, max(n_total, 1), chunk):
        sub = points[i:i + chunk]
        if sub.shape[0] == 0:
            continue
        from . import pipeline
        p = pmvo._pts(sub)
        o, l, hc = pipeline.forward_stage(pmvo, p)                                     # point-sharded under torchrun
        for lst, t in zip(outs, (p, o, l, hc)):
            lst.append(t)
    select_points = torch.cat(outs[0], 0).cpu().numpy()
    select_ori = torch.cat(outs[1], 0).cpu().numpy()
    min_loss = torch.cat(outs[2], 0).cpu().numpy()
    high_conf_index = torch.cat(outs[3], 0).cpu().numpy()
    if _rank0():
        os.makedirs(args.save_root, exist_ok=True)
        np.save(args.save_root + '/select_p.npy', select_points)
        np.save(args.save_root + '/select_o.npy', select_ori)
        np.save(args.save_root + '/min_loss.npy', min_loss)
        np.save(args.save_root + '/high_conf_index.npy', high_conf_index)
    _barrier()
    return select_points, select_ori, min_loss, high_conf_index      # (the reference returns nothing; the files are the contract)


def knn(ref, query, k, dev, cell_factor=1.15):
    """Exact kNN (float64 distances) of `query` [n,3] (float32 or float64) among `ref` [m,3] float32, device tensors."""
    m, n = ref.size(0), query.size(0)
    lo_t, hi_t = ref.amin(0), ref.amax(0)
    lo, hi = lo_t.double().cpu().numpy(), hi_t.double().cpu().numpy()
    ext = np.maximum(hi - lo, 1e-9)
    # Cell size ~ the k-NN radius.  The clouds here are thin shells, so the density is estimated from the cells that
    # are actually occupied at a probe resolution (not from the bounding box): rho = m / (occupied cells * h0^3),
    # r_k = (3k / (4 pi rho))^(1/3).  Too small only costs more shells, too large more candidates; both stay exact.
    h0 = float(max(ext.max() / 128.0, 1e-6))
    sub = ref[:: max(1, m // 200000)]
    cid = ((sub - lo_t) / h0).floor().long()
    ncell = torch.unique(cid[:, 0] * (1 << 40) + cid[:, 1] * (1 << 20) + cid[:, 2]).numel()
    rho = m / max(ncell * h0 ** 3, 1e-30)
    r_k = (3.0 * k / (4.0 * math.pi * rho)) ** (1.0 / 3.0)
    # cell_factor: 1.15 suits queries drawn from the reference cloud itself; queries OFF the cloud (the near-surface points
    # against the selected surface points) reach further and do better with smaller cells (0.65: 14.1 -> 12.1 ms for that
    # stage at BASELINE scale; 0.5-0.9 is flat).  Either way the result is the exact kNN.  env: tuning only
    cell = float(min(max(float(os.environ.get("MH_KNN_CELL_FACTOR", cell_factor)) * r_k, ext.max() / 1000.0), ext.max()))
    bbox = np.concatenate([lo, hi]).astype(np.float64)
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    wsb = lib().mh_knn_workspace_bytes(m, n, k)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        fn = lib().mh_knn_q64 if query.dtype == torch.float64 else lib().mh_knn
        check(fn(stream_ptr(dev), ptr(ref), m, ptr(query), n, k, bbox.ctypes.data_as(C.c_void_p), cell,
                 ptr(idx), ptr(ws), wsb), "mh_knn")
    return idx


def medoid_gather(ori, nbr, dev):
    n, K = nbr.shape
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().mh_medoid_gather(stream_ptr(dev), ptr(ori), ptr(nbr), n, K, ptr(out), None), "mh_medoid_gather")
    return out


_FUSE_PLANES = {}          # (device, grid) -> persistent int32 plane of mh_voxel_fuse (all-zero between calls)


def fuse_plane(dev, grid):
    dev = torch.device(dev)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    gx, gy, gz = [int(g) for g in grid]
    key = (dev.index, gx, gy, gz)
    pl = _FUSE_PLANES.get(key)
    if pl is None:
        pl = torch.empty((lib().mh_voxel_fuse_plane_bytes(gx, gy, gz),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_voxel_fuse_plane_init(stream_ptr(dev), ptr(pl), gx, gy, gz), "mh_voxel_fuse_plane_init")
        _FUSE_PLANES[key] = pl
    return key, pl


def voxel_fuse(select_points, select_ori, dev, grid=GRID, voxel_min=VOXEL_MIN, voxel_size=VOXEL_SIZE, return_index=False,
               valid=None):
    """PMVO.py:695-726 on the device -> float4 volume [gz,gy,gx,4] (see include/monohair_b200.h).  `valid` (optional
    bool/uint8 [n] device tensor): points with a zero entry are skipped, which saves the caller a compaction (and
    the host synchronisation it implies) right before the fusion."""
    pts = torch.as_tensor(select_points).to(dev).type(torch.float).contiguous()
    dirs = torch.as_tensor(select_ori).to(dev).type(torch.float).contiguous()
    n = pts.size(0)
    gx, gy, gz = [int(g) for g in grid]
    vol = torch.empty((gz, gy, gx, 4), dtype=torch.float32, device=dev)
    vidx = torch.empty((n,), dtype=torch.int32, device=dev) if return_index else None
    wsb = lib().mh_voxel_fuse_workspace_bytes(n, gx, gy, gz)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    vmin = np.ascontiguousarray(np.asarray(voxel_min, dtype=np.float64))
    key, plane = fuse_plane(pts.device, grid)
    if valid is not None:
        valid = valid.to(dev).to(torch.uint8).contiguous()
        assert valid.numel() == n
    try:
        with torch.cuda.device(dev):
            check(lib().mh_voxel_fuse(stream_ptr(dev), ptr(pts), ptr(dirs), ptr(valid), n, vmin.ctypes.data_as(C.c_void_p),
                                      float(voxel_size), gx, gy, gz, ptr(vol), ptr(vidx), ptr(plane), ptr(ws), wsb),
                  "mh_voxel_fuse")
    except Exception:
        _FUSE_PLANES.pop(key, None)          # the plane may be dirty: a fresh one is zeroed on the next call
        raise
    return (vol, vidx) if return_index else vol


FUSE_HDR_BYTES = 0          # the persistent fusion plane has no header: all of it is zero between calls


def voxel_fuse_winners(select_points, select_ori, dev, grid=GRID, voxel_min=VOXEL_MIN, voxel_size=VOXEL_SIZE, valid=None,
                       capacity=None):
    """The fusion up to the per-voxel winners (mh_voxel_fuse_winners): -> (winners float32 [capacity,4], count int32 [1])
    on the device; entries past count carry key -1.  Multi-GPU exchange format of the fused volume."""
    pts = torch.as_tensor(select_points).to(dev).type(torch.float).contiguous()
    dirs = torch.as_tensor(select_ori).to(dev).type(torch.float).contiguous()
    n = pts.size(0)
    gx, gy, gz = [int(g) for g in grid]
    cap = int(capacity) if capacity is not None else max(1, min(n, gx * gy * gz))
    win = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    wsb = lib().mh_voxel_fuse_workspace_bytes(n, gx, gy, gz)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    vmin = np.ascontiguousarray(np.asarray(voxel_min, dtype=np.float64))
    key, plane = fuse_plane(pts.device, grid)
    if valid is not None:
        valid = valid.to(dev).to(torch.uint8).contiguous()
    try:
        with torch.cuda.device(dev):
            check(lib().mh_voxel_fuse_winners(stream_ptr(dev), ptr(pts), ptr(dirs), ptr(valid), n,
                                              vmin.ctypes.data_as(C.c_void_p), float(voxel_size), gx, gy, gz, ptr(win), cap,
                                              ptr(cnt), None, ptr(plane), ptr(ws), wsb), "mh_voxel_fuse_winners")
    except Exception:
        _FUSE_PLANES.pop(key, None)
        raise
    return win, cnt


def voxel_scatter(winners, dev, grid=GRID, volume=None):
    """winners float32 [m,4] -> float4 volume (zero filled first unless `volume` is given)."""
    gx, gy, gz = [int(g) for g in grid]
    zero = volume is None
    if volume is None:
        volume = torch.empty((gz, gy, gx, 4), dtype=torch.float32, device=dev)
    winners = winners.contiguous()
    with torch.cuda.device(dev):
        check(lib().mh_voxel_scatter(stream_ptr(dev), ptr(winners), winners.size(0), gx, gy, gz, ptr(volume), 1 if zero else 0),
              "mh_voxel_scatter")
    return volume


def volume_to_mat(vol):
    """float4 volume -> (Occ [gy,gx,gz], Ori [gy,gx,3*gz]) float64 device tensors, the arrays PMVO.py:763-764 saves."""
    gz, gy, gx, _ = vol.shape
    dev = vol.device
    occ = torch.empty((gy, gx, gz), dtype=torch.float64, device=dev)
    ori = torch.empty((gy, gx, 3 * gz), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().mh_volume_to_mat(stream_ptr(dev), ptr(vol), gx, gy, gz, ptr(occ), ptr(ori)), "mh_volume_to_mat")
    return occ, ori


def refine_points(points, ori, loss, pmvo, sub_num=5000, k=100):
    """PMVO.refine step (i) (PMVO.py:605-641) on the device.  kNN is computed once for all points (positions do
    not change); the medoid / re-scoring / update then walks the 5000-point chunks IN ORDER and in place, so
    later chunks gather the orientations already updated by earlier ones, as in the reference (§9-R7)."""
    dev = pmvo.device
    pts = torch.as_tensor(points).to(dev).type(torch.float).contiguous()
    o = torch.as_tensor(ori).to(dev).type(torch.float).contiguous().clone()
    l = torch.as_tensor(loss).to(dev).type(torch.float).contiguous().clone()
    from . import pipeline
    o, l = pipeline.refine_stage(pmvo, pts, o, l, sub_num=sub_num, k=k)
    return pts, o, l


def refine(points, ori, loss, pmvo, filter_unvisible_points, args, infer_inner=True, threshold=0.001,
           genrate_ori_only=False, return_volume=False):
    """PMVO.py:602-764.  Same files written (refine/*.npy, Ori3D.mat / Occ3D.mat in args.save_path)."""
    import scipy.io
    dev = pmvo.device
    if not genrate_ori_only:
        print('filter nosiy points...')
        p_d, o_d, l_d = refine_points(points, ori, loss, pmvo)
        if _rank0():
            os.makedirs(args.output_path + '/refine', exist_ok=True)
            np.save(args.output_path + '/refine/select_p.npy', np.asarray(points))
            np.save(args.output_path + '/refine/select_o.npy', o_d.cpu().numpy())
            np.save(args.output_path + '/refine/min_loss.npy', l_d.cpu().numpy())
        _barrier()                                   # every rank re-reads rank 0's files, like the reference re-reads its own

    points = np.load(args.output_path + '/refine/select_p.npy')
    ori = np.load(args.output_path + '/refine/select_o.npy')
    min_loss = np.load(args.output_path + '/refine/min_loss.npy')
    index = np.where(min_loss < threshold)[0]
    select_ori = torch.from_numpy(ori[index]).to(dev).type(torch.float).contiguous()
    select_points = torch.from_numpy(points[index]).to(dev).type(torch.float).contiguous()

    print('compute points orientation near the surface... ')
    fu_raw = torch.from_numpy(np.ascontiguousarray(filter_unvisible_points)).to(dev).contiguous()
    fu = fu_raw.type(torch.float).contiguous()
    if fu.size(0) > 0 and select_points.size(0) >= 100:
        # the reference queries the KDTree with the points as loaded (float64) and casts them to float32 afterwards (:670-671)
        nbr = knn(select_points, fu_raw if fu_raw.dtype == torch.float64 else fu, 100, dev, cell_factor=0.65)
        filt = pmvo.filter_head_points(fu, args.PMVO.visible_threshold)
        center = medoid_gather(select_ori, nbr, dev)
        fu_ori = center[~filt]
        fu_pts = fu[~filt]
    else:
        fu_ori = torch.zeros((0, 3), device=dev)
        fu_pts = torch.zeros((0, 3), device=dev)
    if _rank0():
        np.save(args.output_path + '/refine/filter_unvisible.npy', fu_pts.cpu().numpy())
        np.save(args.output_path + '/refine/filter_unvisible_ori.npy', fu_ori.cpu().numpy())

    all_ori = torch.cat([select_ori, fu_ori], 0)
    all_pts = torch.cat([select_points, fu_pts], 0)
    vol = voxel_fuse(all_pts, all_ori, dev)

    if infer_inner:
        coarse = np.load(args.data.root + '/ours/raw.npy')
        cp = coarse[:, :3].astype(np.float32)
        co = coarse[:, 3:6].astype(np.float32)
        co[co[:, 1] > 0] *= -1
        cpd = torch.from_numpy(cp).to(dev)
        unvis = pmvo.compute_unvisible_points(cpd)
        up = cpd[unvis].contiguous()
        uo = torch.from_numpy(co).to(dev)[unvis].contiguous()
        gx, gy, gz = GRID
        ws = torch.empty((gx * gy * gz,), dtype=torch.int32, device=dev)
        vmin = np.ascontiguousarray(VOXEL_MIN)
        with torch.cuda.device(dev):
            check(lib().mh_voxel_overwrite(stream_ptr(dev), ptr(up), ptr(uo), up.size(0), vmin.ctypes.data_as(C.c_void_p),
                                           float(VOXEL_SIZE), gx, gy, gz, ptr(vol), ptr(ws)), "mh_voxel_overwrite")
        if _rank0():
            np.save(os.path.join(args.save_path, 'coarse.npy'), up.cpu().numpy())
            np.save(os.path.join(args.save_path, 'coarse_ori.npy'), uo.cpu().numpy())

    if _rank0():
        occ, ori_m = volume_to_mat(vol)
        path = args.save_path
        scipy.io.savemat(os.path.join(path, 'Ori3D.mat'), {'Ori': ori_m.cpu().numpy()})
        scipy.io.savemat(os.path.join(path, 'Occ3D.mat'), {'Occ': occ.cpu().numpy()})
    _barrier()
    if return_volume:
        return vol


def config_parser():
    """PMVO.py:767-800."""
    from . import options
    opt_cmd = options.parse_arguments(sys.argv[1:])
    args = options.set(opt_cmd=opt_cmd)
    args.output_path = os.path.join(args.data.root, args.data.case, args.output_root, args.name)
    os.makedirs(args.output_path, exist_ok=True)
    if int(os.environ.get("RANK", "0")) == 0:
        options.save_options_file(args)
    args.data.root = os.path.join(args.data.root, args.data.case)
    args.bbox_min = np.array(args.bbox_min)
    args.bust_to_origin = np.array(args.bust_to_origin)
    for key in ("strands_path", "bust_path", "raw_points_path", "depth_path", "Ori2D_path", "Conf_path", "mask_path"):
        args.data[key] = os.path.join(args.data.root, args.data[key])
    args.image_camera_path = os.path.join(args.data.root, args.image_camera_path)
    args.save_root = os.path.join(args.output_path, 'optimize')
    if args.PMVO.infer_inner and not args.PMVO.optimize:
        args.save_path = os.path.join(args.output_path, 'full')
    else:
        args.save_path = os.path.join(args.output_path, 'refine')
    os.makedirs(args.save_path, exist_ok=True)
    return args
