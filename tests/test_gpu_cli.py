"""GPU: the drop-in entry points end to end on a tiny capture written in the reference's formats:
`python PMVO.py --yaml=...` then `python HairGrow.py --yaml=...` (same YAML keys, same output files)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pmvo_then_hairgrow_cli(tmp_path):
    from dataset_util import write_capture
    from monohair_b200 import synthetic as syn
    from monohair_b200.hairgrow import load_strand
    sc = syn.make_scene(V=24, H=180, W=240, seed=6)
    write_capture(str(tmp_path / "data"), sc, case="synth")
    cfgdir = tmp_path / "cfg"
    cfgdir.mkdir()
    (cfgdir / "synth.yaml").write_text(f"""_parent_: {ROOT}/configs/reconstruct/base.yaml
name: run1
data:
  root: {tmp_path}/data
  case: synth
  image_size: [180, 240]
PMVO:
  num_sample_per_grid: 2
  patch_size: 5
  threshold: 0.05
  conf_threshold: 0.15
  infer_inner:
HairGenerate:
  grow_threshold: 0.85
  connect_threshold: 0.0025
  connect_dot_threshold: 0.8
  out_ratio: 0.35
""")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "PMVO.py"), f"--yaml={cfgdir}/synth"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    out = tmp_path / "data" / "synth" / "output" / "run1"
    for f in ("optimize/surface.npy", "optimize/filter_unvisible.npy", "optimize/select_p.npy", "optimize/select_o.npy",
              "optimize/min_loss.npy", "optimize/high_conf_index.npy", "refine/select_p.npy", "refine/select_o.npy",
              "refine/min_loss.npy", "refine/filter_unvisible.npy", "refine/filter_unvisible_ori.npy", "refine/Ori3D.mat",
              "refine/Occ3D.mat", "options.yaml"):
        assert (out / f).exists(), f
    sel_o = np.load(out / "optimize/select_o.npy")
    assert sel_o.dtype == np.float32 and sel_o.shape[1] == 3 and sel_o.shape[0] > 1000
    assert np.load(out / "optimize/high_conf_index.npy").dtype == np.bool_
    Occ = scipy.io.loadmat(out / "refine/Occ3D.mat")["Occ"]
    Ori = scipy.io.loadmat(out / "refine/Ori3D.mat")["Ori"]
    assert Occ.shape == (256, 256, 192) and Ori.shape == (256, 256, 576) and Occ.dtype == np.float64
    assert Occ.sum() > 500
    # orientation follows the synthetic flow field
    nz = np.argwhere(Occ > 0)
    o = np.stack([Ori[nz[:, 0], nz[:, 1], nz[:, 2] + c * 192] for c in range(3)], 1)
    centres = (np.stack([nz[:, 1], nz[:, 0], nz[:, 2]], 1) * 0.0025 + np.array([-0.32, -0.32, -0.24])) * np.array([1, -1, -1])
    import torch
    T = syn.flow_tangent(torch.from_numpy(centres), syn.RADII).numpy()
    assert np.median(np.abs((T * o).sum(1))) > 0.97 and np.all(o[:, 1] <= 0)

    r = subprocess.run([sys.executable, os.path.join(ROOT, "HairGrow.py"), f"--yaml={cfgdir}/synth"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    seg, pts = load_strand(str(out / "refine/scalp_segment.hair"))
    num_root = int(np.load(out / "refine/num_root.npy"))
    assert len(seg) > 100 and min(seg) >= 2 and max(seg) <= 513 and pts.shape == (sum(seg), 3)
    assert 0 <= num_root <= len(seg)
    seg_s, pts_s = load_strand(str(out / "refine/scalp_segment_smooth.hair"))       # HairGrow.py:914-916
    assert list(seg_s) == list(seg) and pts_s.shape == pts.shape
    assert 0 < np.abs(pts_s - pts).max() < 0.05                                      # smoothed, not moved away
    # strands live near the shell (world frame, bust offset removed again by VoxelToWorld)
    k = np.linalg.norm((pts + np.array([0.006, -1.644, 0.010])) / np.array(syn.RADII), axis=1)
    assert np.percentile(np.abs(k - 1), 90) < 0.25
    # connect stages (HairGrow.py:925-976): strands.hair keeps every strand (segments possibly extended through their partners),
    # connected_strands.hair keeps the rooted ones
    seg_c, pts_c = load_strand(str(out / "refine/strands.hair"))
    assert len(seg_c) == len(seg) and list(seg_c[:num_root]) == list(seg[:num_root])
    assert all(a >= b for a, b in zip(seg_c, seg)) and sum(seg_c) >= sum(seg)
    seg_f, pts_f = load_strand(str(out / "refine/connected_strands.hair"))
    assert num_root <= len(seg_f) <= len(seg_c) and pts_f.shape == (sum(seg_f), 3) and np.isfinite(pts_f).all()
    k = np.linalg.norm((pts_f + np.array([0.006, -1.644, 0.010])) / np.array(syn.RADII), axis=1)
    assert np.percentile(np.abs(k - 1), 90) < 0.3
