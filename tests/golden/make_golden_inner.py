"""Golden vectors for the infer_inner re-entry of PMVO.refine (PMVO.py:874-880 -> :653-764 with genrate_ori_only=True,
infer_inner=True): the kNN refine is skipped, refine/*.npy are re-voxelised, the invisible points of DeepMVSHair's
raw.npy overwrite the volume (last writer wins), result in full/.  Runs the UNMODIFIED reference on the scene stored in
pmvo_p7.npz.      python tests/golden/make_golden_inner.py      (build container only)"""
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
sys.path.insert(0, os.path.dirname(HERE))

from monohair_b200 import synthetic as syn  # noqa: E402
from golden_util import load, scene_of  # noqa: E402
import ref_import  # noqa: E402


def raw_points(seed=9, m=700):
    """[m,7] = xyz, ori, occ like DeepMVSHair's raw.npy: points from deep inside the shell (hidden in every view) to
    outside it (visible), several per voxel so that the last-writer-wins overwrite matters."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scale = rng.uniform(0.35, 1.15, (m, 1))
    p = d * np.array(syn.RADII) * scale
    p[m // 2:] = p[: m - m // 2] + rng.uniform(-0.0008, 0.0008, (m - m // 2, 3))      # near-duplicates: shared voxels
    o = rng.normal(size=(m, 3))
    o /= np.linalg.norm(o, axis=1, keepdims=True)
    return np.concatenate([p, o, np.ones((m, 1))], 1).astype(np.float64)


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    g = load("pmvo_p7")
    sc = scene_of(g)
    pmvo, mods = ref_import.build_ref_pmvo(sc, patch_size=int(g["patch"]), visible_threshold=1, conf_threshold=float(g["conf_thr"]))
    P = mods["PMVO"]
    from scipy.spatial import KDTree
    scalp = g["scalp"]
    P.device = "cpu"
    P.bust_tree = KDTree(data=scalp)
    P.scalp_tree = KDTree(data=scalp)
    P.scalp_max = np.max(scalp, axis=0)
    raw = raw_points()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(td + "/refine"); os.makedirs(td + "/full"); os.makedirs(td + "/ours")
        np.save(td + "/refine/select_p.npy", g["fwd_points"].astype(np.float32))      # what the first run left behind
        np.save(td + "/refine/select_o.npy", g["ref_select_o"])
        np.save(td + "/refine/min_loss.npy", g["ref_min_loss"])
        np.save(td + "/ours/raw.npy", raw)
        a = types.SimpleNamespace(output_path=td, save_path=td + "/full", device="cpu",
                                  PMVO=types.SimpleNamespace(visible_threshold=1), data=types.SimpleNamespace(root=td))
        P.args = a
        P.refine(None, None, None, pmvo, g["filter_unvisible_in"].copy(), a, infer_inner=True, threshold=float(g["thr"]),
                 genrate_ori_only=True)
        Ori = scipy.io.loadmat(td + "/full/Ori3D.mat")["Ori"]
        Occ = scipy.io.loadmat(td + "/full/Occ3D.mat")["Occ"]
        nz = np.argwhere(Occ > 0)
        Z = Occ.shape[2]
        np.savez_compressed(os.path.join(HERE, "pmvo_p7_inner.npz"), raw=raw, coarse=np.load(td + "/full/coarse.npy"),
                            coarse_ori=np.load(td + "/full/coarse_ori.npy"), mat_occ_nz=nz.astype(np.int32),
                            mat_ori_nz=np.stack([Ori[i, j, [k, k + Z, k + 2 * Z]] for i, j, k in nz]))
        print("pmvo_p7_inner: raw", raw.shape[0], "invisible", np.load(td + "/full/coarse.npy").shape[0], "occupied voxels", len(nz),
              "(without the merge:", len(g["mat_occ_nz"]), ")")


if __name__ == "__main__":
    main()
