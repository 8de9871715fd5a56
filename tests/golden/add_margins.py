"""Record decision margins next to the committed PMVO goldens (build container only):
    python tests/golden/add_margins.py
For every golden written by make_golden.py / make_golden_inner.py this runs the CPU oracle (oracle/margins.py; the
oracle itself is pinned bit-for-bit to the reference's outputs by tests/test_oracle_golden.py) on the golden's own
inputs and stores, per item, how far the reference's discrete choice was from flipping:
  fwd_margin [n]            PMVO.forward: winning (base view, depth sample) loss vs the runner-up
  fwd_thr_gap [n]           distance of the confidence tests behind that choice from their threshold (PMVO.py:196-205)
  fwd_singleton [n]         the reference sampled this point in a batch of one (MKL matrix-vector path; oracle/margins.py)
  ref_knn_gap / ref_medoid_gap / ref_update_gap [n], ref_nbr [n,100]   refine step (i)
  fu_knn_gap / fu_medoid_gap [m]                                        near-surface orientations, step (iii)
  vox_round_gap [p], vox_keys / vox_medoid_gap [q]                      voxelisation
The GPU tests assert exact agreement on every item whose margins exceed their eps and print how many were excluded."""
import os
import sys

import numpy as np
import torch
from scipy.spatial import KDTree

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_util import load, scene_of  # noqa: E402
from oracle import margins as M  # noqa: E402
from oracle import pmvo_oracle as O  # noqa: E402


def pmvo_margins(g, vm, fwd_chunk=None):
    """g: dict-like with the arrays of a PMVO golden -> dict of margin arrays.  fwd_chunk: how many points each
    forward() call of the reference saw when the golden was made (default: all at once)."""
    P, ct, thr = int(g["patch"]), float(g["conf_thr"]), float(g["thr"])
    fm, tg = M.forward_margins(vm, g["fwd_points"], P, ct, with_threshold_gap=True)
    ch = int(fwd_chunk or len(g["fwd_points"]))
    out = {"fwd_margin": fm, "fwd_thr_gap": tg, "fwd_chunk": np.int64(ch),
           "fwd_singleton": M.singleton_base_groups(vm, g["fwd_points"], P, ch)}
    pts32 = g["fwd_points"].astype(np.float32)
    r = M.refine_margins(pts32, g["fwd_ori"], g["ref_select_o"])
    out.update(ref_knn_gap=r["knn_gap"], ref_medoid_gap=r["medoid_gap"], ref_update_gap=r["update_gap"],
               ref_nbr=r["nbr"].astype(np.int32))
    idx = np.where(g["ref_min_loss"] < thr)[0]
    sp, so = pts32[idx], g["ref_select_o"][idx]
    fu = g["filter_unvisible_in"]
    if len(fu) and len(sp) >= 100:
        nn, gap = M.knn_gaps(KDTree(data=sp), fu, 100)
        out.update(fu_knn_gap=gap, fu_medoid_gap=M._medoid_gap(torch.from_numpy(so[nn])))
    else:
        out.update(fu_knn_gap=np.zeros(0), fu_medoid_gap=np.zeros(0))
    allp = np.concatenate([sp, g["ref_fu_points"]]).astype(np.float32)
    allo = np.concatenate([so, g["ref_fu_ori"]]).astype(np.float32)
    rg, keys, mg = M.fuse_margins(allp, allo)
    out.update(vox_round_gap=rg, vox_keys=keys, vox_medoid_gap=mg)
    return out


def main():
    for name in ("pmvo_p7", "pmvo_p5_ties"):
        g = dict(load(name))
        vm = O.ViewMaps.from_scene(scene_of(g))
        g.update(pmvo_margins(g, vm))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
        print(name, "fwd margin min/1e-6-count", float(g["fwd_margin"].min()), int((g["fwd_margin"] < 1e-6).sum()),
              "| refine medoid gap min", float(g["ref_medoid_gap"].min()), "knn gap min", float(g["ref_knn_gap"].min()),
              "| voxel medoid gap min", float(g["vox_medoid_gap"].min()), "round gap min", float(g["vox_round_gap"].min()))


if __name__ == "__main__":
    main()
