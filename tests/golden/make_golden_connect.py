"""Golden vectors for the HairGrow connect stages from the UNMODIFIED reference (build container only, a few minutes):
    python tests/golden/make_golden_connect.py
find_connect_info + connect_segments (HairGrow.py:303-546) and connect_to_scalp (:606-784) hard-code the 256 x 256 x 192
grid at 2.5 mm, so the volume is full size with the hair shell kept only on a cap (a few thousand occupied voxels).
Input strands: guide strands + segments traced by the CPU oracle (pinned bit-for-bit to the reference's trace by
hairgrow_small.npz), pushed through the reference's own VoxelToWorld / .hair round trip exactly as its __main__ does
(:919-976).  numpy's global RNG (retry perturbations, :531) is seeded; the CUDA path seeds it the same way.
The reference runs on the CPU, where torch.from_numpy aliases the strand and points_to_voxel flips it in place on every
retry; the authors ran on CUDA where it does not: torch.from_numpy is patched to copy."""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.io
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from monohair_b200 import synthetic as syn  # noqa: E402
import ref_import  # noqa: E402
from oracle import hairgrow_oracle as H  # noqa: E402
from oracle import pmvo_oracle as O  # noqa: E402

BUST = np.array([0.006, -1.644, 0.010])
THR, DOT_THR, OUT_RATIO, GROW_THR = 0.0025, 0.8, 0.35, 0.85


def cap_volume():
    occ, ori = syn.orientation_volume(shell_mm=3.0)                     # [X,Y,Z], [X,Y,Z,3], world signs
    gx, gy, gz = occ.shape
    iy = np.arange(gy)[None, :, None]
    world_y = -(iy * syn.COARSE_VSIZE + syn.BBOX_MIN[1])
    keep = world_y > 0.085
    occ = occ * keep
    ori = ori * keep[..., None]
    return occ, ori


def scalp_roots(n, seed):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(40 * n, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d[d[:, 1] > 0.75][:n]
    r = np.array(syn.RADII) * 0.93
    p = d * r
    nrm = p / (r * r)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    flip = np.array([1.0, -1.0, -1.0])
    pv = (p * flip - syn.BBOX_MIN) / syn.COARSE_VSIZE
    return pv.astype(np.float32), (nrm * flip).astype(np.float32)


def main():
    mods = ref_import.import_reference()
    HG, U = mods["HairGrow"], mods["Utils"]
    occ, ori = cap_volume()
    M = int((occ > 0).sum())
    print("occupied voxels", M, flush=True)
    vol = H.Volume.from_memory(occ, ori)
    roots, normals = scalp_roots(400, 3)
    rng = np.random.default_rng(11)
    jitter = rng.random((2 * M, 3)).astype(np.float32)
    strands_v, num_root = H.generate_guide_strands(vol, roots, normals, GROW_THR, jitter, passes=2)
    print("strands", len(strands_v), "roots", num_root, flush=True)
    nz = np.argwhere(occ > 0).astype(np.int16)
    out = dict(occ_nz=nz, ori_nz=ori[occ > 0].astype(np.float32), num_root=num_root, bust=BUST, thr=THR, dot_thr=DOT_THR,
               out_ratio=OUT_RATIO, in_len=np.array([s.shape[0] for s in strands_v], np.int32),
               in_pts=np.concatenate(strands_v, 0).astype(np.float32))
    mo, mori = O.mat_layout(occ, ori)
    with tempfile.TemporaryDirectory() as td:
        scipy.io.savemat(td + "/Occ3D.mat", {"Occ": mo})
        scipy.io.savemat(td + "/Ori3D.mat", {"Ori": mori})
        del mo, mori
        solver = HG.HairGrowing(td + "/Occ3D.mat", td + "/Ori3D.mat", device="cpu")
        # ---- the reference's __main__ flow (HairGrow.py:909-976)
        world = solver.VoxelToWorld([torch.from_numpy(s.copy()) for s in strands_v], BUST)
        U.save_hair_strands(td + "/scalp_segment.hair", world)
        segment, points = U.load_strand(td + "/scalp_segment.hair", return_strands=False)
        strands, beg = [], 0
        for i, seg in enumerate(segment):
            s = points[beg:beg + seg]
            if i >= num_root:
                s += BUST
            strands.append(s)
            beg += seg
        solver.strands = strands
        real_from_numpy = torch.from_numpy
        torch.from_numpy = lambda a: real_from_numpy(a.copy())            # CUDA semantics of .to(device): a copy (§9-R16)
        try:
            np.random.seed(123)
            connected = solver.find_connect_info(strands[num_root:], THR, DOT_THR, solver.occ)
        finally:
            torch.from_numpy = real_from_numpy
        new_strands = strands[:num_root] + [c - BUST for c in connected]
        out.update(a_len=np.array([s.shape[0] for s in new_strands], np.int32), a_pts=np.concatenate(new_strands, 0))
        print("stage A: strands", len(new_strands), "points", out["a_pts"].shape[0], flush=True)
        new_strands = U.smooth_strands([s.copy() for s in new_strands], 4.0, 2.0)
        U.save_hair_strands(td + "/strands.hair", new_strands)
        # ---- connect_to_scalp
        segment, points, strands, oris = U.load_strand(td + "/strands.hair", return_strands=True)
        out.update(b_in_len=np.array(segment, np.int32), b_in_pts=points.astype(np.float32))
        HG.args = types.SimpleNamespace(device="cpu", PMVO=types.SimpleNamespace(infer_inner=True),
                                        HairGenerate=types.SimpleNamespace(out_ratio=OUT_RATIO))
        strands = solver.WorldToVoxel(strands, BUST)
        torch.from_numpy = lambda a: real_from_numpy(a.copy())
        try:
            np.random.seed(321)
            cs = solver.connect_to_scalp(strands, num_root)
        finally:
            torch.from_numpy = real_from_numpy
        out.update(b_len=np.array([s.shape[0] for s in cs], np.int32), b_pts=np.concatenate(cs, 0))
        print("stage B: strands", len(cs), "points", out["b_pts"].shape[0], flush=True)
    np.savez_compressed(os.path.join(HERE, "connect_small.npz"), **out)


if __name__ == "__main__":
    main()
