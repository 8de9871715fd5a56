"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py [pmvo|hairgrow|gabor|all]
Writes tests/golden/*.npz (committed).  torch/numpy/scipy versions are recorded in each file.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from monohair_b200 import synthetic as syn  # noqa: E402
import ref_import  # noqa: E402


def versions():
    import scipy
    return np.array([f"torch={torch.__version__}", f"numpy={np.__version__}", f"scipy={scipy.__version__}"])


def scene_arrays(sc):
    return dict(H=sc.H, W=sc.W, poses=np.array([c["pose"] for c in sc.cams]),
                ndc_prj=np.array([c["ndc_prj"] for c in sc.cams]), depth=sc.depth, ori_gray=sc.ori_gray,
                conf_u8=sc.conf_u8, mask_u8=sc.mask_u8)


def gen_pmvo(name, V, H, W, seed, patch, conf_thr, thr, n_cells, n_forward, ori_noise_deg):
    torch.manual_seed(0)
    np.random.seed(0)
    sc = syn.make_scene(V=V, H=H, W=W, seed=seed, ori_noise_deg=ori_noise_deg)
    pmvo, mods = ref_import.build_ref_pmvo(sc, patch_size=patch, visible_threshold=1, conf_threshold=conf_thr)
    P = mods["PMVO"]
    pts = syn.candidate_points(n_cells=n_cells, num_per_grid=2, seed=seed)
    out = dict(versions=versions(), patch=patch, conf_thr=conf_thr, thr=thr, points=pts, **scene_arrays(sc))

    # ---- a10 filter_points through the reference driver (chunking quirk R5 included)
    args = types.SimpleNamespace(device="cpu")
    surface_idx, surface_pts, filter_idx = P.filter_negative_points(pts, pmvo, args)
    n_cov = surface_idx.shape[0]                     # R5: the driver may drop a tail
    out.update(surface_index=surface_idx, filter_index=filter_idx, n_covered=n_cov)

    # ---- a15 forward on the first n_forward surface points
    sel = surface_pts[:n_forward].astype(np.float64)
    p, o, l, hc = pmvo.forward(sel)
    out.update(fwd_points=sel, fwd_ori=o.numpy(), fwd_loss=l.numpy(), fwd_hc=hc.numpy())
    bidx, bval = pmvo.Find_max_conf_from_visible_view()
    out.update(fwd_base_idx=bidx.numpy(), fwd_base_val=bval.numpy(), fwd_visible=pmvo.visible.numpy())

    # ---- a17-a21 refine() end to end, in a temp dir
    scalp = syn.scalp_vertices(800, seed=seed)
    from scipy.spatial import KDTree
    P.device = "cpu"
    P.bust_tree = KDTree(data=scalp)                 # its query result is unused by the reference
    P.scalp_tree = KDTree(data=scalp)
    P.scalp_max = np.max(scalp, axis=0)
    filt_unvis = pts[:n_cov][filter_idx][: n_forward]
    with tempfile.TemporaryDirectory() as td:
        a = types.SimpleNamespace(output_path=td, save_path=os.path.join(td, "refine"), device="cpu",
                                  PMVO=types.SimpleNamespace(visible_threshold=1),
                                  data=types.SimpleNamespace(root=td))
        os.makedirs(a.save_path, exist_ok=True)
        P.args = a
        P.refine(p.numpy().copy(), o.numpy().copy(), l.numpy().copy(), pmvo, filt_unvis.copy(), a,
                 infer_inner=False, threshold=thr, genrate_ori_only=False)
        import scipy.io
        out.update(scalp=scalp, filter_unvisible_in=filt_unvis,
                   ref_select_o=np.load(td + "/refine/select_o.npy"),
                   ref_min_loss=np.load(td + "/refine/min_loss.npy"),
                   ref_fu_points=np.load(td + "/refine/filter_unvisible.npy"),
                   ref_fu_ori=np.load(td + "/refine/filter_unvisible_ori.npy"))
        Ori = scipy.io.loadmat(td + "/refine/Ori3D.mat")["Ori"]
        Occ = scipy.io.loadmat(td + "/refine/Occ3D.mat")["Occ"]
        nz = np.argwhere(Occ > 0)
        out.update(mat_occ_nz=nz.astype(np.int32), mat_ori_shape=np.array(Ori.shape),
                   mat_ori_nz=np.stack([Ori[i, j, [k, k + Occ.shape[2], k + 2 * Occ.shape[2]]] for i, j, k in nz])
                   if len(nz) else np.zeros((0, 3)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "surface", int(surface_idx.sum()), "filter", int(filter_idx.sum()),
          "fwd loss median", float(np.median(l.numpy())), "occ voxels", len(nz))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("pmvo", "all"):
        gen_pmvo("pmvo_p7", V=24, H=135, W=240, seed=1, patch=7, conf_thr=0.15, thr=0.025, n_cells=2500,
                 n_forward=320, ori_noise_deg=4.0)
        gen_pmvo("pmvo_p5_ties", V=22, H=120, W=200, seed=2, patch=5, conf_thr=0.4, thr=0.05, n_cells=1500,
                 n_forward=200, ori_noise_deg=0.0)
    if what in ("hairgrow", "all"):
        import make_golden_hairgrow
        make_golden_hairgrow.main()
    if what in ("gabor", "all"):
        import make_golden_gabor
        make_golden_gabor.main()


if __name__ == "__main__":
    main()
