"""Seeded synthetic capture for the PMVO / HairGrow hot path.

An ellipsoid "wig" seen from a ring of OpenGL-style cameras.  Depth, silhouette,
2-D orientation and confidence maps are analytic and mutually consistent, and
are quantised to the reference's on-disk formats:

* depth  float32 ``(-z_cam/2)*255``, background 255   (Render_utils.py:338-340)
* ori    uint8 gray ``g`` with ``o=(180-g)/180*pi``   (PMVO_utils.py:265-270)
* conf   uint8 ``/255``                               (PMVO_utils.py:272)
* mask   uint8, ``<50 -> 0``, ``/255``                (PMVO_utils.py:302-304)

Used by tests (small sizes, CPU), by ``tests/golden/make_golden.py`` (fed to the
unmodified reference) and by ``bench.py`` (BASELINE.json sizes, on the GPU).
Nothing here is on the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# reference constants: configs/reconstruct/base.yaml:32-34, PMVO.py:695-700
BBOX_MIN = np.array([-0.32, -0.32, -0.24])
COARSE_GRID = (256, 256, 192)
COARSE_VSIZE = 0.005 / 2
FINE_GRID = (512, 512, 384)
FINE_VSIZE = 0.005 / 4

RADII = (0.10, 0.13, 0.11)


@dataclass
class Scene:
    H: int
    W: int
    cams: list                    # [{"file", "pose" (c2w 4x4 list), "ndc_prj" [fx,fy,cx,cy]}]
    depth: np.ndarray             # [V,H,W] float32
    ori_gray: np.ndarray          # [V,H,W] uint8  (best_ori file content)
    conf_u8: np.ndarray           # [V,H,W] uint8
    mask_u8: np.ndarray           # [V,H,W] uint8
    radii: tuple = RADII
    meta: dict = field(default_factory=dict)

    @property
    def V(self):
        return len(self.cams)

    # ---- reference loader equivalents (what PMVO.__init__ receives) ----------
    def ref_depths(self):
        """load_depth output: {view: float32 [H,W,3]} (PMVO_utils.py:278-295)."""
        return {c["file"]: np.repeat(self.depth[i][..., None], 3, axis=-1) for i, c in enumerate(self.cams)}

    def ref_ori_conf(self):
        """Load_Ori_And_Conf output: float64 [H,W,2] / [H,W] (PMVO_utils.py:255-276)."""
        Ori, Conf = {}, {}
        for i, c in enumerate(self.cams):
            o = (180 - self.ori_gray[i]) / 180 * math.pi   # uint8 -> python-int promotion as in the reference
            Ori[c["file"]] = np.stack([np.sin(o), np.cos(o)], -1)
            Conf[c["file"]] = self.conf_u8[i] / 255.
        return Ori, Conf

    def ref_masks(self):
        """load_mask output: float64 [H,W,3] (PMVO_utils.py:297-313)."""
        out = {}
        for i, c in enumerate(self.cams):
            m = np.repeat(self.mask_u8[i][..., None], 3, axis=-1).copy()
            m[m < 50] = 0
            out[c["file"]] = m / 255.
        return out


def _look_at_pose(eye, target, up=(0.0, 1.0, 0.0)):
    """camera-to-world, OpenGL convention (camera looks down -z, +y up)."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(target, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = r, u, -f, eye
    return pose


def make_cameras(V, H, W, radius=0.8, elev_deg=15.0, focal_px=None):
    """Ring of V cameras; ndc_prj=[2f/W, 2f/H, 0, 0] (ingp_utils.py:360)."""
    if focal_px is None:
        focal_px = 1671.0 * (max(H, W) / 1920.0)
    cams = []
    for i in range(V):
        az = 2 * math.pi * i / V
        el = math.radians(elev_deg) * math.sin(3 * az + 0.3)
        eye = [radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az)]
        pose = _look_at_pose(eye, [0.0, 0.0, 0.0])
        cams.append({"file": "view_%03d" % i, "pose": pose.tolist(),
                     "ndc_prj": [2 * focal_px / W, 2 * focal_px / H, 0.0, 0.0]})
    return cams


def flow_tangent(X, radii):
    """Helical flow on the ellipsoid shell: unit tangent at points X [...,3] (torch)."""
    r = torch.tensor(radii, dtype=X.dtype, device=X.device)
    n = X / (r * r)
    n = n / torch.linalg.norm(n, dim=-1, keepdim=True)
    down = torch.tensor([0.0, -1.0, 0.0], dtype=X.dtype, device=X.device).expand_as(X)
    swirl = torch.linalg.cross(n, down, dim=-1)
    t = down + 0.6 * torch.sin(18.0 * X[..., 1:2] + 3.0 * X[..., 0:1]) * swirl + 0.35 * swirl
    t = t - (t * n).sum(-1, keepdim=True) * n
    nrm = torch.linalg.norm(t, dim=-1, keepdim=True)
    fallback = torch.tensor([1.0, 0.0, 0.0], dtype=X.dtype, device=X.device).expand_as(X)
    t = torch.where(nrm > 1e-6, t / nrm.clamp_min(1e-12), fallback)
    return t


def make_scene(V=24, H=270, W=480, seed=0, device="cpu", ori_noise_deg=4.0, radii=RADII,
               radius=0.8, focal_px=None, views=None) -> Scene:
    """`views` (optional range): fill only those views' maps (a rank of a view-sharded upload needs only its block; the
    random stream is consumed for every view, so a view looks the same whoever generates it); the other views' maps are
    left uninitialised."""
    dev = torch.device(device)
    g = torch.Generator().manual_seed(seed)
    cams = make_cameras(V, H, W, radius=radius, focal_px=focal_px)
    rad = torch.tensor(radii, dtype=torch.float64, device=dev)
    depth = np.empty((V, H, W), np.float32)
    ori_gray = np.empty((V, H, W), np.uint8)
    conf_u8 = np.empty((V, H, W), np.uint8)
    mask_u8 = np.empty((V, H, W), np.uint8)
    rows = torch.arange(H, dtype=torch.float64, device=dev)[:, None].expand(H, W)
    cols = torch.arange(W, dtype=torch.float64, device=dev)[None, :].expand(H, W)
    for i, c in enumerate(cams):
        if views is not None and i not in views:
            torch.randn((H, W), generator=g, dtype=torch.float64)
            torch.rand((H, W), generator=g, dtype=torch.float64)
            continue
        fx, fy, cx, cy = c["ndc_prj"]
        pose = torch.tensor(c["pose"], dtype=torch.float64, device=dev)      # c2w
        R, eye = pose[:3, :3], pose[:3, 3]
        # pixel -> camera ray.  x_pix=(-u+1)/2*W, row=(v+1)/2*H with u=fx*x/z+cx, z<0  (PMVO.py:380-382)
        u = -(cols / W * 2 - 1)
        v = rows / H * 2 - 1
        dcam = torch.stack([-(u - cx) / fx, -(v - cy) / fy, -torch.ones_like(u)], -1)   # direction with z=-1
        dw = dcam @ R.T
        # ray / ellipsoid intersection in the scaled space
        o_s, d_s = eye / rad, dw / rad
        a = (d_s * d_s).sum(-1)
        b = 2 * (d_s * o_s).sum(-1)
        cc = (o_s * o_s).sum() - 1.0
        disc = b * b - 4 * a * cc
        hit = disc > 0
        t = (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a)                    # depth along -z_cam (dcam.z=-1)
        X = eye + t[..., None] * dw
        dmap = torch.where(hit, t / 2 * 255, torch.full_like(t, 255.0))
        # projected tangent -> (d_row, d_col)
        T = flow_tangent(X, radii)
        Xc = (X - eye) @ R                                                   # world -> camera (R orthonormal)
        Tc = T @ R
        eps = 1e-3
        P1 = Xc + eps * Tc

        def pix(Pc):
            uu = fx * Pc[..., 0] / Pc[..., 2] + cx
            vv = fy * Pc[..., 1] / Pc[..., 2] + cy
            return (vv + 1) / 2 * H, (-uu + 1) / 2 * W
        r0, c0 = pix(Xc)
        r1, c1 = pix(P1)
        ang = torch.atan2(r1 - r0, c1 - c0)                                  # (sin o, cos o) = (d_row, d_col)
        noise = torch.randn((H, W), generator=g, dtype=torch.float64).to(dev) * math.radians(ori_noise_deg)
        ang = torch.remainder(ang + noise, math.pi)
        gray = torch.remainder(torch.round(180.0 - ang * 180.0 / math.pi), 180.0)
        cnoise = torch.rand((H, W), generator=g, dtype=torch.float64).to(dev)
        # confidence: higher where the view is frontal, with per-pixel noise; < 0.1 off hair
        n_w = X / (rad * rad)
        n_w = n_w / torch.linalg.norm(n_w, dim=-1, keepdim=True).clamp_min(1e-12)
        facing = (-(n_w * dw).sum(-1) / torch.linalg.norm(dw, dim=-1)).clamp(0, 1)
        conf = torch.where(hit, (0.12 + 0.55 * facing + 0.45 * cnoise * facing + 0.08 * cnoise).clamp(0, 1),
                           0.08 * cnoise)
        depth[i] = dmap.to(torch.float32).cpu().numpy()
        ori_gray[i] = torch.where(hit, gray, torch.zeros_like(gray)).to(torch.uint8).cpu().numpy()
        conf_u8[i] = torch.round(conf * 255).to(torch.uint8).cpu().numpy()
        mask_u8[i] = torch.where(hit, 255, 0).to(torch.uint8).cpu().numpy()
    return Scene(H=H, W=W, cams=cams, depth=depth, ori_gray=ori_gray, conf_u8=conf_u8, mask_u8=mask_u8,
                 radii=tuple(radii), meta={"seed": seed, "ori_noise_deg": ori_noise_deg})


def candidate_points(n_cells=None, num_per_grid=4, shell_mm=3.0, seed=0, radii=RADII, vsize=FINE_VSIZE,
                     grid=FINE_GRID, bbox_min=BBOX_MIN, max_points=None):
    """What load_colmap_points / SamplePointsAroundmesh return for this shape
    (PMVO_utils.py:316-362): occupied cells of the fine grid within ``shell_mm`` of
    the surface, ``num_per_grid`` uniform points per cell, cells in np.nonzero
    (x,y,z) order, the whole cell list repeated num_per_grid times.  float64 [N,3]."""
    rng = np.random.default_rng(seed)
    r = np.asarray(radii)
    ext = (r + shell_mm * 1e-3 + 2 * vsize)
    lo = np.floor((np.array([-ext[0], -ext[1], -ext[2]]) - bbox_min) / vsize).astype(int)
    hi = np.ceil((np.array([ext[0], ext[1], ext[2]]) - bbox_min) / vsize).astype(int)
    lo = np.maximum(lo, 0)
    hi = np.minimum(hi, np.array(grid) - 1)
    ix, iy, iz = np.meshgrid(np.arange(lo[0], hi[0] + 1), np.arange(lo[1], hi[1] + 1),
                             np.arange(lo[2], hi[2] + 1), indexing="ij")
    # voxel space is (x,-y,-z): cell centre in flipped space
    c = np.stack([ix, iy, iz], -1).reshape(-1, 3) * vsize + bbox_min
    q = c / r                                       # flip of y,z does not change the ellipsoid
    k = np.linalg.norm(q, axis=-1)
    # first-order distance to the surface
    grad = np.linalg.norm(c / (r * r), axis=-1) / np.maximum(k, 1e-12)
    dist = (k - 1.0) / np.maximum(grad, 1e-12)
    cells = np.stack([ix, iy, iz], -1).reshape(-1, 3)[np.abs(dist) < shell_mm * 1e-3]
    if n_cells is not None and cells.shape[0] > n_cells:
        sel = np.sort(rng.choice(cells.shape[0], n_cells, replace=False))
        cells = cells[sel]
    base = np.concatenate([cells] * num_per_grid, 0).astype(np.float64)
    sample = (base + rng.random(base.shape)) * vsize + bbox_min
    sample[:, 1:] *= -1
    if max_points is not None:
        sample = sample[:max_points]
    return sample


def scalp_vertices(n=2000, seed=0, radii=RADII, scale=0.92):
    """Upper-cap vertices of a slightly smaller ellipsoid: stands in for ours/scalp_tsfm.obj."""
    rng = np.random.default_rng(seed + 17)
    d = rng.normal(size=(n * 3, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d[d[:, 1] > 0.15][:n]
    return d * np.asarray(radii) * scale


def orientation_volume(grid=COARSE_GRID, vsize=COARSE_VSIZE, bbox_min=BBOX_MIN, radii=RADII,
                       shell_mm=4.0, device="cpu"):
    """Analytic Occ/Ori volume in the in-memory layout of PMVO.refine (PMVO.py:695-726):
    occ [X,Y,Z] float64, ori [X,Y,Z,3] float64 (world frame, ori.y<=0)."""
    dev = torch.device(device)
    gx, gy, gz = grid
    ix = torch.arange(gx, dtype=torch.float64, device=dev)
    iy = torch.arange(gy, dtype=torch.float64, device=dev)
    iz = torch.arange(gz, dtype=torch.float64, device=dev)
    X = torch.stack(torch.meshgrid(ix, iy, iz, indexing="ij"), -1) * vsize + torch.tensor(bbox_min, device=dev)
    X = X * torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64, device=dev)    # voxel (x,-y,-z) -> world
    r = torch.tensor(radii, dtype=torch.float64, device=dev)
    k = torch.linalg.norm(X / r, dim=-1)
    grad = torch.linalg.norm(X / (r * r), dim=-1) / k.clamp_min(1e-12)
    dist = (k - 1.0) / grad.clamp_min(1e-12)
    occ = (dist.abs() < shell_mm * 1e-3).to(torch.float64)
    T = flow_tangent(X / k.clamp_min(1e-12)[..., None], radii)
    T = torch.where(T[..., 1:2] > 0, -T, T)
    ori = T * occ[..., None]
    return occ.cpu().numpy(), ori.cpu().numpy()
