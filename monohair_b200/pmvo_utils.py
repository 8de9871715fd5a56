"""Host-side mirror of the hot-path parts of Utils/PMVO_utils.py (same names): loaders of the on-disk formats
(SURVEY.md §3.5), candidate-point sampling, p2v, the medoid, .mat readers.  Device work goes through the C ABI.
open3d / trimesh are not required: OBJ meshes are read and sampled with numpy.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import pmvo as _P


# ------------------------------------------------------------------------------------------------ loaders
def Load_Ori_And_Conf(camera, Ori_path, Conf_path):
    """PMVO_utils.py:255-276 (including its suffix probing)."""
    import cv2
    Ori, Conf = {}, {}
    suffix = '.JPG'
    for view, _ in camera.items():
        if not os.path.exists(os.path.join(Ori_path, view + '.JPG')):
            suffix = '.png'
        if not os.path.exists(os.path.join(Ori_path, view + '.png')):
            suffix = '.jpg'
        o = cv2.imread(os.path.join(Ori_path, view + suffix), cv2.IMREAD_GRAYSCALE)
        o = (180 - o) / 180 * math.pi
        Ori[view] = np.stack([np.sin(o), np.cos(o)], -1)
        Conf[view] = cv2.imread(os.path.join(Conf_path, view + suffix), cv2.IMREAD_GRAYSCALE) / 255.
    return Ori, Conf


def load_depth(camera, path, type='npy'):
    """PMVO_utils.py:278-295."""
    return {view: np.load(os.path.join(path, view + '.npy')).astype(np.float32) for view, _ in camera.items()}


def load_mask(camera, path):
    """PMVO_utils.py:297-313."""
    import cv2
    files = os.listdir(path)
    suffix = files[0][-4:]
    masks = {}
    for view, _ in camera.items():
        mask = cv2.imread(os.path.join(path, view + suffix))
        mask[mask < 50] = 0
        masks[view] = mask / 255.
    return masks


def load_u8_maps(camera, Ori_path, Conf_path, mask_path, depth_path):
    """File formats straight to the arrays PMVO.from_u8 consumes (no float64 decode on the host):
    depth float32 [V,H,W], ori gray / conf / mask uint8 [V,H,W]."""
    import cv2
    files = os.listdir(mask_path)
    msuf = files[0][-4:]
    d, o, c, m = [], [], [], []
    suffix = '.JPG'
    for view, _ in camera.items():
        if not os.path.exists(os.path.join(Ori_path, view + '.JPG')):
            suffix = '.png'
        if not os.path.exists(os.path.join(Ori_path, view + '.png')):
            suffix = '.jpg'
        o.append(cv2.imread(os.path.join(Ori_path, view + suffix), cv2.IMREAD_GRAYSCALE))
        c.append(cv2.imread(os.path.join(Conf_path, view + suffix), cv2.IMREAD_GRAYSCALE))
        m.append(cv2.imread(os.path.join(mask_path, view + msuf))[..., 0])
        dd = np.load(os.path.join(depth_path, view + '.npy'))
        d.append(np.ascontiguousarray(dd[..., 0] if dd.ndim == 3 else dd, dtype=np.float32))
    return np.stack(d), np.stack(o), np.stack(c), np.stack(m)


def read_obj(path):
    """minimal Wavefront OBJ reader: vertices [n,3] float64, triangle faces [m,3] int64 (polygons are fanned)."""
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith('v '):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('f '):
                idx = [int(tok.split('/')[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.array(vs, dtype=np.float64).reshape(-1, 3), np.array(fs, dtype=np.int64).reshape(-1, 3)


def read_obj_normals(path):
    """OBJ with per-vertex normals the way open3d's read_triangle_mesh presents them (HairGrow.py:879): the `vn` record
    a face corner refers to becomes that vertex's normal (last corner wins); vertices no corner gives a normal to, or
    files without `vn`, fall back to area-weighted face normals.  -> vertices [n,3], faces [m,3], vertex normals [n,3]."""
    vs, vns, fs, fns = [], [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith('v '):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('vn '):
                vns.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('f '):
                toks = [tok.split('/') for tok in line.split()[1:]]
                idx = [int(t[0]) for t in toks]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                nid = [int(t[2]) if len(t) > 2 and t[2] != '' else 0 for t in toks]
                nid = [i - 1 if i > 0 else (len(vns) + i if i < 0 else -1) for i in nid]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
                    fns.append([nid[0], nid[k], nid[k + 1]])
    v = np.array(vs, dtype=np.float64).reshape(-1, 3)
    fa = np.array(fs, dtype=np.int64).reshape(-1, 3)
    a, b, c = v[fa[:, 0]], v[fa[:, 1]], v[fa[:, 2]]
    fn = np.cross(b - a, c - a)                                  # length = 2 x area: area weighting for free
    vn = np.zeros_like(v)
    for k in range(3):
        np.add.at(vn, fa[:, k], fn)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    if vns:
        table = np.array(vns, dtype=np.float64).reshape(-1, 3)
        fna = np.array(fns, dtype=np.int64).reshape(-1, 3)
        ok = fna >= 0
        vn[fa[ok]] = table[fna[ok]]                              # numpy scatter: the last corner written wins
    return v, fa, vn


def sample_points_uniformly(vertices, faces, number_of_points, rng=None, with_normals=False, vertex_normals=None):
    """area-weighted uniform surface sampling (stands in for open3d's sample_points_uniformly, whose RNG is
    unseeded in the reference, PMVO_utils.py:346 / HairGrow.py:881).  With `vertex_normals` the returned normals are
    the barycentric blend of the triangle's vertex normals with the sampling weights, un-normalised -- what open3d
    returns for use_triangle_normal=False (HairGrow.py:880-881); otherwise flat face normals."""
    # default: a generator seeded from numpy's global state, which options.process_options seeds (seed: 0), so a run is
    # repeatable (open3d's own sampler is unseeded in the reference)
    rng = np.random.default_rng(int(np.random.randint(0, 2 ** 31 - 1))) if rng is None else rng
    a, b, c = vertices[faces[:, 0]], vertices[faces[:, 1]], vertices[faces[:, 2]]
    n = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(n, axis=1)
    tri = rng.choice(len(faces), size=number_of_points, p=area / area.sum())
    r1, r2 = np.sqrt(rng.random(number_of_points)), rng.random(number_of_points)
    pts = (1 - r1)[:, None] * a[tri] + (r1 * (1 - r2))[:, None] * b[tri] + (r1 * r2)[:, None] * c[tri]
    if with_normals and vertex_normals is not None:
        n0, n1, n2 = vertex_normals[faces[tri, 0]], vertex_normals[faces[tri, 1]], vertex_normals[faces[tri, 2]]
        return pts, (1 - r1)[:, None] * n0 + (r1 * (1 - r2))[:, None] * n1 + (r1 * r2)[:, None] * n2
    if with_normals:
        nn = n[tri] / np.maximum(np.linalg.norm(n[tri], axis=1, keepdims=True), 1e-20)
        return pts, nn
    return pts


def SamplePointsAroundmesh(colmap_points, bbox_min, vsize, num_per_grid=32, grid_resolution=[512, 512, 384], device=None):
    """PMVO_utils.py:316-339 (np.random.random, seeded through options.process_options).  With a CUDA `device` the cell
    marking, the compaction in np.nonzero order and the sample arithmetic run on the device (csrc/sample.cu, float64, the
    reference's operation order); the random numbers are still drawn from numpy's global stream on the host, in the same
    shape and order, so the result is bit-identical to the host path.  Like the reference, flips colmap_points in place."""
    if device is not None and torch.device(device).type == "cuda":
        import ctypes as C
        from ._lib import check, lib, ptr, stream_ptr
        dev = torch.device(device)
        gx, gy, gz = [int(g) for g in grid_resolution]
        pts = torch.from_numpy(np.ascontiguousarray(colmap_points, dtype=np.float64)).to(dev)
        colmap_points[:, 1:] *= -1                                 # the in-place flip the caller observes (:318)
        bmin = np.ascontiguousarray(np.asarray(bbox_min, dtype=np.float64))
        occ = torch.empty((gx, gy, gz), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_sample_mark_cells(stream_ptr(dev), ptr(pts), pts.size(0), bmin.ctypes.data_as(C.c_void_p), float(vsize),
                                             gx, gy, gz, ptr(occ)), "mh_sample_mark_cells")
        cells = torch.nonzero(occ).contiguous()                    # lexicographic (x, y, z), like np.nonzero
        m = int(cells.size(0))
        rnd = torch.from_numpy(np.random.random((m * num_per_grid, 3))).to(dev)
        out = torch.empty((m * num_per_grid, 3), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_sample_cells(stream_ptr(dev), ptr(cells), m, int(num_per_grid), ptr(rnd), bmin.ctypes.data_as(C.c_void_p),
                                        float(vsize), ptr(out)), "mh_sample_cells")
        return out.cpu().numpy()
    occ = np.zeros(grid_resolution, dtype=bool)
    colmap_points[:, 1:] *= -1
    indexs = np.round((colmap_points - bbox_min) / vsize).astype(np.int32)
    x = np.clip(indexs[:, 0], 0, grid_resolution[0] - 1)
    y = np.clip(indexs[:, 1], 0, grid_resolution[1] - 1)
    z = np.clip(indexs[:, 2], 0, grid_resolution[2] - 1)
    occ[x, y, z] = True
    x, y, z = np.nonzero(occ)
    indices = np.concatenate([x[:, None], y[:, None], z[:, None]], 1)
    base = np.concatenate([indices] * num_per_grid, 0)
    sample = (base + np.random.random(base.shape[:]) * 1) * vsize + bbox_min
    sample[:, 1:] *= -1
    return sample


def load_colmap_points(path, bbox_min, bust_to_origin, vsize=0.005, grid_resolution=[128, 128, 96], sample=True, num_per_grid=8,
                       device=None):
    """PMVO_utils.py:341-362 (`device`: run the cell sampling on that CUDA device)."""
    v, f = read_obj(path)
    print('num_p:', v.shape[0])
    colmap_points = sample_points_uniformly(v, f, v.shape[0] * 5)
    colmap_points += bust_to_origin
    if sample:
        sample_points = SamplePointsAroundmesh(colmap_points.copy(), bbox_min, vsize, num_per_grid=num_per_grid,
                                               grid_resolution=grid_resolution, device=device)
        print('num sample:', sample_points.shape[:])
        return sample_points
    return colmap_points


def load_bust(path):
    v, f = read_obj(path)
    return v, f, None


# ------------------------------------------------------------------------------------------------ math
def compute_points_similarity(ori):
    """PMVO_utils.py:366-382: medoid of ori [N,K,3] (device tensor) under |cos| -> [N,3]."""
    N, K, _ = ori.shape
    dev = ori.device
    flat = ori.reshape(N * K, 3).type(torch.float).contiguous()
    nbr = torch.arange(N * K, dtype=torch.int32, device=dev).reshape(N, K).contiguous()
    return _P.medoid_gather(flat, nbr, dev)


def p2v(points, voxel_min, voxel_size, grid_resolution):
    """PMVO_utils.py:386-404 (host, float64; flips points[:,1:] in place like the reference, §9-R6).  The device
    version lives in csrc/voxel_fuse.cu."""
    points[:, 1:] *= -1
    indexs = np.round((points - voxel_min) / voxel_size).astype(np.int32)
    x = np.clip(indexs[:, 0], 0, grid_resolution[0] - 1)
    y = np.clip(indexs[:, 1], 0, grid_resolution[1] - 1)
    z = np.clip(indexs[:, 2], 0, grid_resolution[2] - 1)
    return x, y, z


def voxel_to_points(voxels):
    from .hairgrow import voxel_to_points as v2p
    return v2p(voxels)


def points_to_voxel(points):
    from .hairgrow import points_to_voxel as p2vox
    return p2vox(points)


def get_ground_truth_3D_occ(d, flip=False):
    """PMVO_utils.py:86-95 -> [Z,Y,X,1] float32."""
    import scipy.io
    occ = scipy.io.loadmat(d, verify_compressed_data_integrity=False)['Occ'].astype(np.float32)
    occ = np.expand_dims(np.transpose(occ, [2, 0, 1]), -1)
    if flip:
        occ = occ[:, :, ::-1, :]
    return np.ascontiguousarray(occ)


def get_ground_truth_3D_ori(d, flip=False, growInv=False):
    """PMVO_utils.py:98-113 -> [Z,Y,X,3] float32."""
    import scipy.io
    ori = scipy.io.loadmat(d, verify_compressed_data_integrity=False)['Ori'].astype(np.float32)
    ori = np.reshape(ori, [ori.shape[0], ori.shape[1], 3, -1]).transpose([0, 1, 3, 2]).transpose(2, 0, 1, 3)
    if flip:
        ori = ori[:, :, ::-1, :] * np.array([-1.0, 1.0, 1.0])
    return np.ascontiguousarray(ori)
