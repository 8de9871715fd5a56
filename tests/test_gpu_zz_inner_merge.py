"""GPU: the infer_inner re-entry of PMVO.refine (what infer_inner.py:89-90 triggers with `--PMVO.infer_inner
--PMVO.optimize=`): stored refine results re-voxelised, invisible points of raw.npy written over the volume (last writer
wins), files in full/ -- against goldens from the unmodified reference (tests/golden/make_golden_inner.py)."""
import os
import types

import numpy as np
import pytest
import scipy.io
from scipy.spatial import KDTree

from golden_util import load, scene_of

pytestmark = pytest.mark.gpu


def test_inner_merge_vs_reference_golden(tmp_path):
    from monohair_b200 import pmvo as P
    from monohair_b200.camera import cameras_from_scene
    g, gi = load("pmvo_p7"), load("pmvo_p7_inner")
    sc = scene_of(g)
    Ori, Conf = sc.ref_ori_conf()
    pm = P.PMVO(cameras_from_scene(sc), sc.ref_depths(), Ori, Conf, sc.ref_masks(), device="cuda:0",
                image_size=[sc.H, sc.W], patch_size=int(g["patch"]), visible_threshold=1, conf_threshold=float(g["conf_thr"]))
    scalp = g["scalp"]
    P.scalp_tree = KDTree(data=scalp)
    P.scalp_max = scalp.max(0)
    td = str(tmp_path)
    for d in ("refine", "full", "ours"):
        os.makedirs(os.path.join(td, d), exist_ok=True)
    np.save(td + "/refine/select_p.npy", g["fwd_points"].astype(np.float32))       # what the first PMVO run left behind
    np.save(td + "/refine/select_o.npy", g["ref_select_o"])
    np.save(td + "/refine/min_loss.npy", g["ref_min_loss"])
    np.save(td + "/ours/raw.npy", gi["raw"])
    a = types.SimpleNamespace(output_path=td, save_path=td + "/full", device="cuda:0",
                              PMVO=types.SimpleNamespace(visible_threshold=1), data=types.SimpleNamespace(root=td))
    P.refine(None, None, None, pm, g["filter_unvisible_in"].copy(), a, infer_inner=True, threshold=float(g["thr"]),
             genrate_ori_only=True)
    assert np.array_equal(np.load(td + "/full/coarse.npy"), gi["coarse"])            # invisible set: exact
    assert np.array_equal(np.load(td + "/full/coarse_ori.npy"), gi["coarse_ori"])
    Occ = scipy.io.loadmat(td + "/full/Occ3D.mat")["Occ"]
    Ori3 = scipy.io.loadmat(td + "/full/Ori3D.mat")["Ori"]
    nz = np.argwhere(Occ > 0)
    assert np.array_equal(nz, gi["mat_occ_nz"])                                       # occupancy after the merge: exact
    Z = Occ.shape[2]
    vals = np.stack([Ori3[i, j, [k, k + Z, k + 2 * Z]] for i, j, k in nz])
    same = np.all(vals == gi["mat_ori_nz"], axis=1)
    print(f"\ninner merge: {same.mean() * 100:.2f}% of {len(nz)} occupied voxels bit-identical")
    # the inputs are the reference's own stored refine results, so no kNN / selection ambiguity enters; the voxel medoid
    # means are reproduced bit for bit (gates.py): every voxel must agree
    assert same.all(), f"{int((~same).sum())} voxels differ"
    # every voxel the merge wrote carries the LAST raw point's orientation that fell into it: exact
    from oracle import pmvo_oracle as O
    x, y, z = O.p2v(gi["coarse"].astype(np.float64).copy(), np.array([-0.32, -0.32, -0.24]), 0.005 / 2, np.array([256, 256, 192]))
    last = {}
    for i, key in enumerate(zip(y.tolist(), x.tolist(), z.tolist())):                # .mat layout: Occ[Y][X][Z]
        last[key] = i
    for (yy, xx, zz), i in last.items():
        assert np.array_equal(Ori3[yy, xx, [zz, zz + Z, zz + 2 * Z]], gi["coarse_ori"][i].astype(np.float64))
