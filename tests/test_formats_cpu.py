"""CPU: the on-disk formats either side of the hot path: loaders against the scene they were written
from, the .hair codec, the OBJ reader / sampler, candidate sampling."""
import math
import os

import numpy as np

from dataset_util import ellipsoid_mesh, write_capture, write_obj
from monohair_b200 import synthetic as syn
from monohair_b200.camera import load_cam, parsing_camera
from monohair_b200 import pmvo_utils as U


def test_loaders_roundtrip(tmp_path):
    sc = syn.make_scene(V=3, H=40, W=56, seed=4)
    d = write_capture(str(tmp_path), sc)
    cams = parsing_camera(load_cam(os.path.join(d, "ours", "cam_params.json")), os.path.join(d, "capture_images"))
    assert list(cams.keys()) == [c["file"] for c in sc.cams]
    c0 = cams["view_000"]
    assert np.allclose(c0.pose.numpy(), np.linalg.inv(np.array(sc.cams[0]["pose"])).astype(np.float32))
    assert float(c0.proj[0, 0]) == np.float32(sc.cams[0]["ndc_prj"][0]) and float(c0.proj[3, 2]) == -1.0
    Ori, Conf = U.Load_Ori_And_Conf(cams, os.path.join(d, "best_ori"), os.path.join(d, "conf"))
    Ori_r, Conf_r = sc.ref_ori_conf()
    masks, depths = U.load_mask(cams, os.path.join(d, "hair_mask")), U.load_depth(cams, os.path.join(d, "render_depth"))
    for k in cams:
        assert Ori[k].dtype == np.float64 and np.array_equal(Ori[k], Ori_r[k]) and np.array_equal(Conf[k], Conf_r[k])
        assert masks[k].shape == (40, 56, 3) and np.array_equal(masks[k], sc.ref_masks()[k])
        assert depths[k].dtype == np.float32 and np.array_equal(depths[k], sc.ref_depths()[k])
    dd, oo, cc, mm = U.load_u8_maps(cams, os.path.join(d, "best_ori"), os.path.join(d, "conf"), os.path.join(d, "hair_mask"),
                                    os.path.join(d, "render_depth"))
    assert np.array_equal(dd, sc.depth) and np.array_equal(oo, sc.ori_gray) and np.array_equal(cc, sc.conf_u8)
    assert np.array_equal(mm, sc.mask_u8)


def test_parsing_camera_subsampling(tmp_path):
    cam = [{"file": "v%04d" % i, "pose": np.eye(4).tolist(), "ndc_prj": [3.0, 1.7, 0, 0]} for i in range(40)]
    img = tmp_path / "imgs"
    img.mkdir()
    for i in range(301):
        (img / ("f%d.png" % i)).write_bytes(b"")
    assert len(parsing_camera(cam, str(img))) == 20            # > 300 files -> every 2nd camera (Camera_utils.py:152-155)
    for i in range(301, 501):
        (img / ("f%d.png" % i)).write_bytes(b"")
    assert len(parsing_camera(cam, str(img))) == 10            # > 500 files -> every 4th
    assert len(parsing_camera(cam, None)) == 40


def test_hair_codec_matches_reference_layout(tmp_path):
    import struct
    from monohair_b200.hairgrow import load_strand, save_hair_strands
    rng = np.random.default_rng(0)
    strands = [rng.normal(size=(n, 3)).astype(np.float32) for n in (5, 17, 513)]
    p = str(tmp_path / "s.hair")
    save_hair_strands(p, strands)
    raw = open(p, "rb").read()
    n_s, n_p = struct.unpack("II", raw[:8])                    # Utils.py:1246-1262
    assert (n_s, n_p) == (3, 535)
    assert struct.unpack("HHH", raw[8:14]) == (5, 17, 513)
    assert len(raw) == 8 + 2 * 3 + 4 * 3 * 535
    seg, pts = load_strand(p)
    assert seg == [5, 17, 513] and np.array_equal(pts.astype(np.float32), np.concatenate(strands))


def test_obj_reader_and_candidate_sampling(tmp_path):
    v, f = ellipsoid_mesh(syn.RADII, 20, 40)
    p = str(tmp_path / "m.obj")
    write_obj(p, v, f)
    v2, f2 = U.read_obj(p)
    assert np.allclose(v2, v, atol=1e-8) and np.array_equal(f2, f)
    pts, nrm = U.sample_points_uniformly(v2, f2, 5000, rng=np.random.default_rng(1), with_normals=True)
    k = np.linalg.norm(pts / np.array(syn.RADII), axis=1)
    assert np.all(np.abs(k - 1) < 0.02) and np.allclose(np.linalg.norm(nrm, axis=1), 1)
    np.random.seed(0)
    s = U.SamplePointsAroundmesh(pts.copy(), np.array([-0.32, -0.32, -0.24]), 0.005 / 4, num_per_grid=2, grid_resolution=[512, 512, 384])
    assert s.shape[0] % 2 == 0 and s.shape[1] == 3
    half = s.shape[0] // 2                                      # the cell list is repeated num_per_grid times (PMVO_utils.py:334)
    cell = lambda a: np.floor((a * np.array([1, -1, -1]) - np.array([-0.32, -0.32, -0.24])) / (0.005 / 4) + 1e-9).astype(int)
    assert np.array_equal(cell(s[:half]), cell(s[half:]))
    x, y, z = U.p2v(s.copy(), np.array([-0.32, -0.32, -0.24]), 0.0025, np.array([256, 256, 192]))
    assert x.min() >= 0 and x.max() < 256 and z.max() < 192
