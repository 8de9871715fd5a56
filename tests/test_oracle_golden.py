"""CPU: the oracle restatement (oracle/pmvo_oracle.py) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  Bit-exact: both run the same torch CPU primitives."""
import numpy as np
import pytest
import torch
from scipy.spatial import KDTree

from golden_util import load, scene_of
from oracle import pmvo_oracle as O

CASES = ["pmvo_p7", "pmvo_p5_ties"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    g = load(request.param)
    sc = scene_of(g)
    return g, O.ViewMaps.from_scene(sc)


def test_filter_points_bit_exact(case):
    g, vm = case
    pts = torch.from_numpy(g["points"][: int(g["n_covered"])]).float()
    s, f, _ = O.filter_points(vm, pts, int(g["patch"]), 1, float(g["conf_thr"]))
    assert np.array_equal(s.numpy(), g["surface_index"])
    assert np.array_equal(f.numpy(), g["filter_index"])


def test_forward_bit_exact(case):
    g, vm = case
    n = 96                                       # per-point independent: a prefix is enough on CPU
    _, o, l, hc = O.forward(vm, g["fwd_points"], int(g["patch"]), float(g["conf_thr"]))
    assert np.array_equal(o.numpy()[:n], g["fwd_ori"][:n])
    assert np.array_equal(l.numpy(), g["fwd_loss"])
    assert np.array_equal(hc.numpy(), g["fwd_hc"])
    assert np.array_equal(o.numpy(), g["fwd_ori"])


def test_refine_and_voxelise_bit_exact(case):
    g, vm = case
    P, ct, thr = int(g["patch"]), float(g["conf_thr"]), float(g["thr"])
    scalp = g["scalp"]
    tree, smax = KDTree(data=scalp), scalp.max(0)
    p, o, l = O.refine_points(vm, g["fwd_points"].astype(np.float32), g["fwd_ori"], g["fwd_loss"], P, 1, ct, tree, smax)
    assert np.array_equal(o, g["ref_select_o"])
    assert np.array_equal(l, g["ref_min_loss"])
    idx = np.where(l < thr)[0]
    fp, fo = O.unvisible_orientation(vm, p[idx], o[idx], g["filter_unvisible_in"], 1, tree, smax)
    assert np.array_equal(fp, g["ref_fu_points"])
    assert np.array_equal(fo, g["ref_fu_ori"])
    occ, ori = O.voxel_fuse(np.concatenate([p[idx], fp]), np.concatenate([o[idx], fo]))
    mo, mori = O.mat_layout(occ, ori)
    nz = np.argwhere(mo > 0)
    assert np.array_equal(nz, g["mat_occ_nz"])
    Z = mo.shape[2]
    vals = np.stack([mori[i, j, [k, k + Z, k + 2 * Z]] for i, j, k in nz])
    assert np.array_equal(vals, g["mat_ori_nz"])
    assert tuple(mori.shape) == tuple(g["mat_ori_shape"])


def test_inner_merge_bit_exact():
    """infer_inner re-entry (PMVO.py:874-880): re-voxelise the stored refine results, then overwrite with the invisible
    points of raw.npy -- oracle vs the unmodified reference's full/Occ3D.mat, Ori3D.mat, coarse*.npy."""
    g = load("pmvo_p7")
    gi = load("pmvo_p7_inner")
    vm = O.ViewMaps.from_scene(scene_of(g))
    scalp = g["scalp"]
    tree, smax = KDTree(data=scalp), scalp.max(0)
    p, o, l = g["fwd_points"].astype(np.float32), g["ref_select_o"], g["ref_min_loss"]
    idx = np.where(l < float(g["thr"]))[0]
    fp, fo = O.unvisible_orientation(vm, p[idx], o[idx], g["filter_unvisible_in"], 1, tree, smax)
    occ, ori = O.voxel_fuse(np.concatenate([p[idx], fp]), np.concatenate([o[idx], fo]))
    up, uo = O.merge_inner(vm, occ, ori, gi["raw"])
    assert np.array_equal(up, gi["coarse"]) and np.array_equal(uo, gi["coarse_ori"])
    assert 0 < up.shape[0] < gi["raw"].shape[0]
    mo, mori = O.mat_layout(occ, ori)
    nz = np.argwhere(mo > 0)
    assert np.array_equal(nz, gi["mat_occ_nz"]) and len(nz) > len(g["mat_occ_nz"])
    Z = mo.shape[2]
    vals = np.stack([mori[i, j, [k, k + Z, k + 2 * Z]] for i, j, k in nz])
    assert np.array_equal(vals, gi["mat_ori_nz"])
