"""Golden vectors for the HairGrow trace from the UNMODIFIED reference (HairGrow.py) on a small synthetic volume.
torch.rand_like is patched to return pre-drawn rows so the oracle / CUDA path can consume identical jitter."""
import os
import sys
import tempfile

import numpy as np
import scipy.io
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from monohair_b200 import synthetic as syn  # noqa: E402
import ref_import  # noqa: E402

GRID = (48, 48, 40)
VSIZE = 0.64 / 48


def small_volume():
    occ, ori = syn.orientation_volume(grid=GRID, vsize=VSIZE, shell_mm=1.3 * VSIZE * 1e3)
    return occ, ori


def scalp_roots(n, seed):
    """roots on a smaller ellipsoid's upper cap, in VOXEL coordinates, with voxel-frame normals."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(4 * n, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d[d[:, 1] > 0.2][:n]
    r = np.array(syn.RADII) * 0.8
    p = d * r
    nrm = p / (r * r)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    flip = np.array([1.0, -1.0, -1.0])
    pv = (p * flip - syn.BBOX_MIN) / VSIZE
    return pv.astype(np.float32), (nrm * flip).astype(np.float32)


def main():
    mods = ref_import.import_reference()
    HG = mods["HairGrow"]
    occ, ori = small_volume()
    import oracle.pmvo_oracle as O
    mo, mori = O.mat_layout(occ, ori)
    roots, normals = scalp_roots(60, 0)
    M = int((occ > 0).sum())
    rng = np.random.default_rng(7)
    jitter = rng.random((3 * M, 3)).astype(np.float32)
    out = dict(occ=occ.astype(np.uint8), ori=ori.astype(np.float32), roots=roots, normals=normals, jitter=jitter,
               grid=np.array(GRID), thr=0.85)
    with tempfile.TemporaryDirectory() as td:
        scipy.io.savemat(td + "/Occ3D.mat", {"Occ": mo})
        scipy.io.savemat(td + "/Ori3D.mat", {"Ori": mori})
        for name, fn_name, passes in (("guide", "GenerateGuideStrandFromScalp", 2), ("segments", "randomlyGenerateSegments", 3)):
            solver = HG.HairGrowing(td + "/Occ3D.mat", td + "/Ori3D.mat", device="cpu")
            state = {"i": 0}
            orig = torch.rand_like

            def fake_rand_like(t, *a, **k):
                r = torch.from_numpy(jitter[state["i"]].copy())
                state["i"] += 1
                return r
            torch.rand_like = fake_rand_like
            try:
                if name == "guide":
                    strands, num_root = solver.GenerateGuideStrandFromScalp(torch.from_numpy(roots), torch.from_numpy(normals),
                                                                            None, 0.85)
                else:
                    strands, num_root = solver.randomlyGenerateSegments(0.85), 0
            finally:
                torch.rand_like = orig
            assert state["i"] == passes * M, (state["i"], passes, M)
            lens = np.array([s.shape[0] for s in strands], np.int32)
            pts = torch.cat(strands, 0).numpy() if len(strands) else np.zeros((0, 3), np.float32)
            out[name + "_len"] = lens
            out[name + "_pts"] = pts
            out[name + "_num_root"] = num_root
            print(name, "strands", len(strands), "roots", num_root, "points", pts.shape[0], "M", M)
    np.savez_compressed(os.path.join(HERE, "hairgrow_small.npz"), **out)


if __name__ == "__main__":
    main()
