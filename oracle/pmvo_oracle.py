"""TEST INFRASTRUCTURE — CPU oracle for the PMVO hot path.  NOT part of the product.

A restatement, in torch-CPU / numpy, of the reference's PMVO algorithm
(/root/reference/PMVO.py, Utils/Camera_utils.py, Utils/PMVO_utils.py), each
function citing the reference lines it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the UNMODIFIED reference, run in the build
container on seeded synthetic scenes by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/*.npz`` (``tests/test_oracle_golden.py`` checks them
bit-for-bit where the reference is deterministic).

It uses the same torch CPU primitives as the reference where the primitive's
arithmetic matters (``torch.matmul`` projection, ``torch.topk`` tie order,
``torch.sum`` cascade order, ``torch.round`` half-even), so its float results are
bit-identical to the reference's on the same host.
"""
from __future__ import annotations

import math

import numpy as np
import torch

ZFAR, ZNEAR = 100.0, 0.1


class ViewMaps:
    """The per-view inputs PMVO.__init__ receives (PMVO.py:14-37), as float32 CPU tensors.

    depth/mask keep only channel 0 (the only channel get_depth / get_mask read, PMVO.py:485,523).
    """

    def __init__(self, cams, depths, Ori, Conf, masks, image_size):
        self.keys = [c["file"] for c in cams]
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.pose, self.proj = [], []
        for c in cams:
            # Camera.__init__ / get_projection_matrix (Camera_utils.py:11-36); pose = inv(c2w) (:160)
            fx, fy, cx, cy = c["ndc_prj"]
            proj = np.array([[fx, 0, cx, 0], [0, fy, cy, 0],
                             [0, 0, (-ZFAR - ZNEAR) / (ZFAR - ZNEAR), -2. * ZFAR * ZNEAR / (ZFAR - ZNEAR)],
                             [0, 0, -1, 0]])
            self.proj.append(torch.from_numpy(proj).type(torch.float))
            self.pose.append(torch.from_numpy(np.linalg.inv(np.array(c["pose"]))).type(torch.float))
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).type(torch.float)
        self.depth = [f32(np.asarray(depths[k])[..., 0] if np.asarray(depths[k]).ndim == 3 else depths[k]) for k in self.keys]
        self.mask = [f32(np.asarray(masks[k])[..., 0] if np.asarray(masks[k]).ndim == 3 else masks[k]) for k in self.keys]
        self.ori = [f32(Ori[k]) for k in self.keys]
        self.conf = [f32(Conf[k]) for k in self.keys]

    @property
    def V(self):
        return len(self.keys)

    @classmethod
    def from_scene(cls, scene):
        Ori, Conf = scene.ref_ori_conf()
        depths = {c["file"]: scene.depth[i] for i, c in enumerate(scene.cams)}
        masks_full = scene.ref_masks()
        masks = {k: v[..., 0] for k, v in masks_full.items()}
        return cls(scene.cams, depths, Ori, Conf, masks, [scene.H, scene.W])


# --------------------------------------------------------------------------- a5
def projection(pose, proj, pts):
    """Camera.projection (Camera_utils.py:38-58): returns ndc uv [N,2] and camera z [N]."""
    hom = torch.cat([pts.permute(1, 0), torch.ones((1, pts.size(0)))])
    cam = torch.matmul(pose, hom)
    z = cam[2:3, :]
    uv = torch.matmul(proj, cam)
    uv[:2] /= z
    return uv.transpose(1, 0)[:, :2], z[0]


def ndc_to_xy(uv, H, W):
    """PMVO.py:380-382 / uv2pixel before the flip (Camera_utils.py:67-69): float (x_pix, y_pix)."""
    uv = uv.clone()
    uv[:, 0:1] = -uv[:, 0:1]
    uv[:, :2] = (uv[:, :2] + 1) / 2
    uv[:, :2] *= torch.tensor([W, H], dtype=torch.float)
    return uv


def reprojection_world(pose, proj, uv, z):
    """Camera.reprojection(..., to_world=True) (Camera_utils.py:81-106).

    The reference fills a [M,4] row-major buffer viewed as [4,M] and multiplies by the (column-major)
    output of torch.linalg.inv; MKL's accumulation order for this 3x3 product depends on those layouts
    (and on M), so the same layouts are reproduced here to stay bit-identical on the same host."""
    cam = torch.empty((uv.size(0), 4)).permute(1, 0)
    cam[0] = (uv[:, 0] - proj[0, 2]) / proj[0, 0] * z
    cam[1] = (uv[:, 1] - proj[1, 2]) / proj[1, 1] * z
    cam[2] = z
    cam[3] = 1
    world = torch.matmul(torch.linalg.inv(pose[:3, :3]), cam[:3] - pose[:3, 3:4])
    return world.permute(1, 0)


# --------------------------------------------------------------------------- a6
def project_points(vm: ViewMaps, v, pts):
    """PMVO.project_points (PMVO.py:378-397): (row,col) int64, z=-z_cam/2, out-of-image flag."""
    uv, z = projection(vm.pose[v], vm.proj[v], pts)
    xy = torch.round(ndc_to_xy(uv, vm.H, vm.W)).type(torch.long)
    oob = (xy[:, 0] > vm.W - 1) | (xy[:, 0] < 0) | (xy[:, 1] > vm.H - 1) | (xy[:, 1] < 0)
    col = torch.clamp(xy[:, 0], 0, vm.W - 1)
    row = torch.clamp(xy[:, 1], 0, vm.H - 1)
    return row, col, -z / 2, oob


# --------------------------------------------------------------------------- a7
def patch_offsets(P):
    """get_ori_patch / get_c_patch scan order (PMVO.py:494-500, 507-511): row offset outer, col inner."""
    h = P // 2
    return [(di, dj) for di in range(-h, h + 1) for dj in range(-h, h + 1)]


def gather_patch(m, row, col, P):
    H, W = m.shape[0], m.shape[1]
    out = []
    for di, dj in patch_offsets(P):
        out.append(m[torch.clamp(row + di, 0, H - 1), torch.clamp(col + dj, 0, W - 1)])
    return torch.stack(out, 1)          # [N,P*P] or [N,P*P,2]


def compute_visible(depth, z255):
    """PMVO.compute_visible (PMVO.py:525-529)."""
    d = z255 - depth
    vis = torch.where(d < 0.1, 1 - d / 0.1, torch.full_like(d, -1.0))
    return torch.clamp(vis, -1, 1)


# --------------------------------------------------------------------------- a9
def compute_visible_and_ori(vm: ViewMaps, pts, P):
    """PMVO.Compute_Visible_and_Ori (PMVO.py:346-376).  The unused `mask` stack is omitted."""
    vis, ori, conf, op, cp = [], [], [], [], []
    for v in range(vm.V):
        row, col, z, oob = project_points(vm, v, pts)
        vb = compute_visible(vm.depth[v][row, col], z * 255.)
        vb[oob] = -1
        vis.append(vb)
        ori.append(vm.ori[v][row, col])
        conf.append(vm.conf[v][row, col])
        op.append(gather_patch(vm.ori[v], row, col, P))
        cp.append(gather_patch(vm.conf[v], row, col, P))
    return {"visible": torch.stack(vis), "Ori": torch.stack(ori),
            "Conf": torch.clamp(torch.stack(conf), 1e-6, 1),
            "Ori_patch": torch.stack(op), "Conf_patch": torch.clamp(torch.stack(cp), 1e-6, 1)}


# --------------------------------------------------------------------------- a10
def filter_points(vm: ViewMaps, pts, P, visible_threshold, conf_threshold):
    """PMVO.filter_points (PMVO.py:402-459) -> (surface bool[N], filter bool[N], counters float[5,N])."""
    N = pts.size(0)
    s_vis = torch.zeros(N)
    s_vism = torch.zeros(N)
    s_idx = torch.zeros(N)
    s_vis1 = torch.zeros(N)
    s_vis1m = torch.zeros(N)
    rows = {k: [] for k in ("vis", "vism", "idx", "vis1", "vis1m")}
    for v in range(vm.V):
        row, col, z, oob = project_points(vm, v, pts)
        m = vm.mask[v][row, col].clone()
        d = vm.depth[v][row, col]
        c = gather_patch(vm.conf[v], row, col, P).max(dim=-1)[0]
        c[oob] = 0
        delta = z * 255 - d
        unvis = (delta > 0.1).float()
        unvis[oob] = 1
        unvis1 = (delta > visible_threshold).float()
        unvis1[oob] = 1
        low_c = (c < conf_threshold).float()
        m[m > 0.2] = 1
        rows["vis"].append(1 - unvis)
        rows["vism"].append((1 - unvis) * m)
        rows["idx"].append((1 - unvis) * low_c)
        rows["vis1"].append(1 - unvis1)
        rows["vis1m"].append((1 - unvis1) * m)
    # torch.sum(dim=0) over the stacked views, exactly as PMVO.py:442-449
    s_vis, s_vism, s_idx, s_vis1, s_vis1m = (torch.sum(torch.stack(rows[k]), 0)
                                             for k in ("vis", "vism", "idx", "vis1", "vis1m"))
    low = s_idx > 4
    hair = (s_vis - s_vism) < s_vis * 1 / 2
    hair1 = (s_vis1 - s_vis1m) < s_vis1 * 1 / 2
    surface = s_vis > 1
    filt = (s_vis1 > 1) & ~surface
    surface = surface & (~low & hair)
    filt = filt & (~low & hair1)
    return surface, filt, torch.stack([s_vis, s_vism, s_idx, s_vis1, s_vis1m])


def compute_unvisible_points(vm: ViewMaps, pts):
    """PMVO.compute_unvisible_points (PMVO.py:461-480)."""
    cnt = []
    for v in range(vm.V):
        row, col, z, oob = project_points(vm, v, pts)
        d = vm.depth[v][row, col]
        unvis = (z * 255 - d > 0.9).float()
        unvis[oob] = 1
        cnt.append(1 - unvis)
    return ~(torch.sum(torch.stack(cnt), 0) > 2)


# --------------------------------------------------------------------------- a11
def find_base_views(visible, Conf, k=20):
    """PMVO.Find_max_conf_from_visible_view (PMVO.py:339-343).  torch.topk itself is used so the
    (implementation-defined) order among equal values is the reference's on CPU."""
    c = torch.where(visible < 1, Conf * torch.maximum(visible, torch.zeros_like(visible)), Conf)
    val, idx = torch.topk(c, k, dim=0, largest=True)
    return idx, val, c


def sample_offsets(num_sample=90):
    """PMVO.py:274-278."""
    s1 = torch.arange(-0.005, -0.001, 0.004 / (num_sample / 4))
    s2 = torch.arange(-0.001, 0.001, 0.002 / (num_sample / 2))
    s3 = torch.arange(0.001, 0.005, 0.004 / (num_sample / 4))
    return torch.cat([s1, s2, s3], 0)[:num_sample]


# --------------------------------------------------------------------------- a12
def sample_next_3d_pos(vm: ViewMaps, pts, base_view, Ori, num_sample=90):
    """PMVO.sample_next_3d_pos (PMVO.py:263-335).  surface_points == points so only
    the sample cloud [N,S,3] is returned."""
    off = sample_offsets(num_sample)
    S = off.numel()
    out = torch.zeros((pts.size(0), S, 3))
    size = torch.tensor([vm.W, vm.H], dtype=torch.float)
    for v in range(vm.V):
        sel = base_view == v
        if int(sel.sum()) == 0:
            continue
        uv, z = projection(vm.pose[v], vm.proj[v], pts[sel])
        xy = ndc_to_xy(uv, vm.H, vm.W)
        nxt = xy + Ori[v][sel][:, [1, 0]] * 2           # (d_row,d_col) -> (dx,dy), 2 px step (:300)
        nxt = nxt / size
        nxt = nxt * 2 - 1
        nxt[:, 0:1] = -nxt[:, 0:1]
        zs = (z[:, None].expand(-1, S) + off).reshape(-1)
        nx = nxt[:, None, :].expand(-1, S, -1).reshape(-1, 2)
        world = reprojection_world(vm.pose[v], vm.proj[v], nx, zs)
        out[sel] = world.reshape(-1, S, 3)
    return out


# --------------------------------------------------------------------------- a13
def to_pixel_rc(vm, v, p):
    """projection + uv2pixel (Camera_utils.py:60-71): float (row, col)."""
    uv, _ = projection(vm.pose[v], vm.proj[v], p)
    return torch.flip(ndc_to_xy(uv, vm.H, vm.W), dims=[1])


def compute_reproject_ori(vm: ViewMaps, pts, samples):
    """PMVO.compute_reproject_ori (PMVO.py:219-241) -> [V,N,S,2] in (d_row,d_col)."""
    N, S = samples.size(0), samples.size(1)
    flat = samples.reshape(-1, 3)
    out = []
    for v in range(vm.V):
        ps = to_pixel_rc(vm, v, flat).reshape(N, S, 2)
        p0 = to_pixel_rc(vm, v, pts)
        out.append(ps - p0[:, None, :])
    return torch.stack(out)


# --------------------------------------------------------------------------- a14
def compute_prj_loss(state, prj, conf_threshold, return_all=False):
    """PMVO.compute_prj_loss (PMVO.py:151-209).  The per-entry update (lines 173-182) is the
    sequential scan of the reference: entry 0 seeds unconditionally; entry p replaces iff
    l_p < loss and (Conf_p > thr if the patch has any entry > thr, else always)."""
    Op, Cp, vis = state["Ori_patch"], state["Conf_patch"], state["visible"]
    S = prj.size(2)
    hi = (Cp.max(-1)[0] > conf_threshold)[..., None]           # [V,N,1]
    loss = best_c = None
    for p in range(Cp.size(-1)):
        o = Op[:, :, p, :][:, :, None, :].expand(-1, -1, S, -1)
        c = Cp[:, :, p][:, :, None].expand(-1, -1, S)
        sim = torch.maximum(torch.cosine_similarity(o, prj, dim=-1), torch.cosine_similarity(-o, prj, dim=-1))
        l = 1 - sim
        if loss is None:
            loss, best_c = l, c
        else:
            lt = l < loss
            take = (lt & (c > conf_threshold) & hi) | (lt & ~hi)
            loss = torch.where(take, l, loss)
            best_c = torch.where(take, c, best_c)
    w = torch.where(vis == -1, torch.zeros_like(vis), torch.ones_like(vis))[:, :, None] * best_c   # compute_weight (:211-215)
    lw = loss * w
    sw = torch.sum(w, dim=0)
    ratio = sw / torch.sum(w > 0, dim=0)
    pos = ratio > conf_threshold                              # [N,S]
    low = torch.sum(pos, dim=-1) < 5                          # [N]
    L = torch.sum(lw, dim=0) / sw
    raw = L.clone()
    L = torch.where(pos, L, torch.ones_like(L))
    L[low] = raw[low]
    mn, am = torch.min(L, dim=-1)
    hc = pos[torch.arange(pos.size(0)), am]
    if return_all == "detail":                                # for oracle/margins.py: the decisions behind L
        return mn, am, hc, L, dict(raw=raw, ratio=ratio, pos=pos, low=low)
    if return_all:
        return mn, am, hc, L
    return mn, am, hc


# --------------------------------------------------------------------------- a15
def forward(vm: ViewMaps, points_np, P, conf_threshold, debug=False):
    """PMVO.forward (PMVO.py:39-78) -> (points, ori, loss, high_conf)."""
    pts = torch.from_numpy(np.asarray(points_np)).type(torch.float)
    st = compute_visible_and_ori(vm, pts, P)
    bidx, bval, _ = find_base_views(st["visible"], st["Conf"])
    best = torch.zeros_like(pts)
    min_loss = hcs = None
    dbg = {"L": [], "base": bidx[0:20:2].clone(), "base_conf": bval[0:20:2].clone(), "arg": [], "loss_b": []}
    for i in range(0, 20, 2):
        smp = sample_next_3d_pos(vm, pts, bidx[i], st["Ori"])
        prj = compute_reproject_ori(vm, pts, smp)
        if debug:
            loss, am, hc, L, det = compute_prj_loss(st, prj, conf_threshold, return_all="detail")
            dbg["L"].append(L)
            dbg.setdefault("detail", []).append(det)
            dbg["arg"].append(am.clone())
            dbg["loss_b"].append(loss.clone())
        else:
            loss, am, hc = compute_prj_loss(st, prj, conf_threshold)
        if min_loss is None:
            min_loss, hcs = loss, hc
            upd = torch.ones(pts.size(0), dtype=torch.bool)
        else:
            upd = (loss < min_loss) & (bval[i] > 0)
            min_loss[upd] = loss[upd]
            hcs[upd] = hc[upd]
        best[upd] = smp[upd, am[upd], :]
    d = best - pts
    ori = d / torch.linalg.norm(d, 2, dim=-1, keepdim=True)
    if debug:
        dbg["best_sample"] = best
        return pts, ori, min_loss, hcs, dbg
    return pts, ori, min_loss, hcs


# --------------------------------------------------------------------------- a17
def filter_head_points(vm: ViewMaps, pts, visible_threshold, scalp_tree, scalp_max):
    """PMVO.filter_head_points (PMVO.py:96-144).  The bust_tree query result is unused there."""
    pn = pts.clone().cpu().numpy()
    dist, _ = scalp_tree.query(pn, k=1)
    head_top = torch.from_numpy(np.logical_and(dist < 0.04, pn[:, 2] < scalp_max[2] - 0.01))
    s_vis, s_idx = [], []
    for v in range(vm.V):
        row, col, z, _ = project_points(vm, v, pts)
        m = vm.mask[v][row, col].clone()
        d = vm.depth[v][row, col]
        unvis = (z * 255 - d >= visible_threshold).float()
        m[m > 0.2] = 1
        s_idx.append((1 - unvis) * m)
        s_vis.append(1 - unvis)
    sv = torch.sum(torch.stack(s_vis), 0)
    si = torch.sum(torch.stack(s_idx), 0)
    return ~((sv - si) < sv * 1 / 2) & ~head_top


# --------------------------------------------------------------------------- a18
def refine_loss(vm: ViewMaps, pts, ori, P, visible_threshold, conf_threshold, scalp_tree, scalp_max):
    """PMVO.refine (PMVO.py:81-93)."""
    st = compute_visible_and_ori(vm, pts, P)
    filt = filter_head_points(vm, pts, visible_threshold, scalp_tree, scalp_max)
    nxt = (pts + ori * 0.005 / 4)[:, None, :]
    prj = compute_reproject_ori(vm, pts, nxt)
    loss, _, _ = compute_prj_loss(st, prj, conf_threshold)
    loss[filt] = -1
    return loss


# --------------------------------------------------------------------------- a19
def compute_points_similarity(ori):
    """PMVO_utils.compute_points_similarity (PMVO_utils.py:366-382): medoid of [N,K,3] under |cos|."""
    N, K, _ = ori.size()
    a = ori[:, :, None, :].expand(-1, -1, K, -1)
    b = a.permute(0, 2, 1, 3)
    sim = torch.maximum(torch.cosine_similarity(a, b, dim=-1), torch.cosine_similarity(-a, b, dim=-1))
    k = torch.argmax(torch.mean(sim, dim=-1), dim=-1)
    return ori[torch.arange(N), k], k


# --------------------------------------------------------------------------- a21
def p2v(points, voxel_min, voxel_size, grid_resolution):
    """PMVO_utils.p2v (PMVO_utils.py:386-404); flips y,z of `points` IN PLACE like the reference (§9-R6)."""
    points[:, 1:] *= -1
    idx = np.round((points - voxel_min) / voxel_size).astype(np.int32)
    x = np.clip(idx[:, 0], 0, grid_resolution[0] - 1)
    y = np.clip(idx[:, 1], 0, grid_resolution[1] - 1)
    z = np.clip(idx[:, 2], 0, grid_resolution[2] - 1)
    return x, y, z


# --------------------------------------------------------------------------- a20 (v)
def voxel_fuse(select_points, select_ori, grid=(256, 256, 192), voxel_min=(-0.32, -0.32, -0.24), voxel_size=0.005 / 2):
    """PMVO.refine, voxelisation part (PMVO.py:695-726): flip to ori.y<=0, p2v, per-voxel medoid.
    -> occ float64 [X,Y,Z], ori float64 [X,Y,Z,3] (in-memory layout before the .mat shuffle)."""
    grid = np.array(grid).astype(np.int32)
    occ = np.zeros(grid)
    ori = np.zeros((*grid, 3))
    so = np.array(select_ori, copy=True)
    so[so[:, 1] > 0] *= -1
    x, y, z = p2v(np.array(select_points, copy=True), np.array(voxel_min), voxel_size, grid)
    key = (x.astype(np.int64) * grid[1] + y) * grid[2] + z
    order = np.argsort(key, kind="stable")            # dict insertion keeps per-voxel point order
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], ks.size]
    for s, e in zip(starts, ends):
        idx = order[s:e]
        val = torch.from_numpy(so[idx]).type(torch.float)
        med, _ = compute_points_similarity(val[None])
        occ[x[idx[0]], y[idx[0]], z[idx[0]]] = 1
        ori[x[idx[0]], y[idx[0]], z[idx[0]]] = med[0].numpy()
    return occ, ori


def merge_inner(vm, occ, ori, raw, grid=(256, 256, 192), voxel_min=(-0.32, -0.32, -0.24), voxel_size=0.005 / 2):
    """PMVO.refine, infer_inner branch (PMVO.py:733-751): the points of DeepMVSHair's raw.npy ([M,7] = xyz, ori, occ)
    that no view sees overwrite the fused volume; orientations flipped to y <= 0; numpy advanced-index assignment, so
    among several points of one voxel the LAST one wins.  occ/ori are modified in place.
    -> (un_visible_points [m,3] float32, unvisible_ori [m,3] float32) = what the reference saves as coarse*.npy."""
    points = raw[:, :3].astype(np.float32)
    c_ori = raw[:, 3:6].astype(np.float32)
    c_ori[c_ori[:, 1] > 0] *= -1
    unv = compute_unvisible_points(vm, torch.from_numpy(points)).numpy()
    up, uo = points[unv], c_ori[unv]
    x, y, z = p2v(up.copy(), np.array(voxel_min), voxel_size, np.array(grid).astype(np.int32))
    occ[x, y, z] = 1
    ori[x, y, z] = uo
    return up, uo


def mat_layout(occ, ori):
    """PMVO.py:753-756: in-memory [X,Y,Z(,3)] -> arrays saved as Occ3D.mat / Ori3D.mat."""
    g = occ.shape
    o = ori.transpose((0, 1, 3, 2)).reshape(g[0], g[1], g[2] * 3).transpose((1, 0, 2))
    return occ.transpose((1, 0, 2)), o


# --------------------------------------------------------------------------- a20 (i)-(iii)
def refine_points(vm, points, ori, loss, P, visible_threshold, conf_threshold, scalp_tree, scalp_max,
                  sub_num=5000, k=100):
    """PMVO.refine module function, step (i) (PMVO.py:605-641): kNN medoid smoothing and re-scoring,
    chunk-sequential and in place."""
    from scipy.spatial import KDTree
    points, ori, loss = np.array(points, copy=True), np.array(ori, copy=True), np.array(loss, copy=True)
    tree = KDTree(data=points)
    for i in range(points.shape[0] // sub_num + 1):
        sl = slice(i * sub_num, min((i + 1) * sub_num, points.shape[0]))
        if points[sl].shape[0] == 0:
            continue
        _, nn = tree.query(points[sl], k)
        center, _ = compute_points_similarity(torch.from_numpy(ori[nn]))
        sp, so = torch.from_numpy(points[sl]), torch.from_numpy(ori[sl])
        upd = refine_loss(vm, sp, center, P, visible_threshold, conf_threshold, scalp_tree, scalp_max)
        sim = torch.maximum(torch.cosine_similarity(center, so, dim=-1), torch.cosine_similarity(center, -so, dim=-1))
        ch = sim < 0.95
        so[ch] = center[ch]
        upd[upd == -1] = 0.5
        ori[sl] = so.numpy()
        loss[sl] = upd.numpy()
    return points, ori, loss


def unvisible_orientation(vm, select_points, select_ori, filter_unvisible_points, visible_threshold,
                          scalp_tree, scalp_max, sub_num=5000, k=100):
    """PMVO.refine step (iii) (PMVO.py:660-686)."""
    from scipy.spatial import KDTree
    tree = KDTree(data=select_points)
    oo, pp = [], []
    for i in range(filter_unvisible_points.shape[0] // sub_num + 1):
        sub = filter_unvisible_points[i * sub_num:min((i + 1) * sub_num, filter_unvisible_points.shape[0])]
        if sub.shape[0] == 0:
            continue
        _, nn = tree.query(sub, k)
        sp = torch.from_numpy(sub).type(torch.float)
        filt = filter_head_points(vm, sp, visible_threshold, scalp_tree, scalp_max)
        center, _ = compute_points_similarity(torch.from_numpy(select_ori[nn]))
        oo.append(center[~filt])
        pp.append(sp[~filt])
    return torch.cat(pp, 0).numpy(), torch.cat(oo, 0).numpy()
