"""Write a synthetic capture to disk in the reference's on-disk formats for CLI / loader tests."""
import json
import os

import numpy as np


def ellipsoid_mesh(radii, n_lat=24, n_lon=48, cap_only=False):
    """lat-long triangulated ellipsoid (y up).  cap_only: upper cap (scalp stand-in)."""
    lat = np.linspace(0.02, (0.42 if cap_only else 0.98) * np.pi, n_lat)
    lon = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    v = np.stack([np.sin(la) * np.sin(lo), np.cos(la), np.sin(la) * np.cos(lo)], -1).reshape(-1, 3) * np.asarray(radii)
    f = []
    for i in range(n_lat - 1):
        for j in range(n_lon):
            a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
            c, d = a + n_lon, b + n_lon
            f += [[a, c, b], [b, c, d]]
    return v, np.array(f)


def write_obj(path, v, f):
    with open(path, "w") as fh:
        for p in v:
            fh.write("v %.9f %.9f %.9f\n" % tuple(p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(t + 1))


def write_capture(root, scene, case="synth", bust_to_origin=(0.006, -1.644, 0.010)):
    """data/<case>/{ours/cam_params.json, capture_images/, render_depth/*.npy, best_ori/*.png, conf/*.png,
    hair_mask/*.png, ours/*.obj}"""
    import cv2
    d = os.path.join(root, case)
    for sub in ("ours", "capture_images", "render_depth", "best_ori", "conf", "hair_mask"):
        os.makedirs(os.path.join(d, sub), exist_ok=True)
    json.dump({"cam_list": [{"file": c["file"], "pose": c["pose"], "ndc_prj": c["ndc_prj"]} for c in scene.cams]},
              open(os.path.join(d, "ours", "cam_params.json"), "w"))
    for i, c in enumerate(scene.cams):
        k = c["file"]
        open(os.path.join(d, "capture_images", k + ".png"), "wb").close()          # only listed, never read by PMVO
        np.save(os.path.join(d, "render_depth", k + ".npy"), np.repeat(scene.depth[i][..., None], 3, -1).astype(np.float32))
        cv2.imwrite(os.path.join(d, "best_ori", k + ".png"), scene.ori_gray[i])
        cv2.imwrite(os.path.join(d, "conf", k + ".png"), scene.conf_u8[i])
        cv2.imwrite(os.path.join(d, "hair_mask", k + ".png"), np.repeat(scene.mask_u8[i][..., None], 3, -1))
    off = np.asarray(bust_to_origin)
    r = np.asarray(scene.radii)
    v, f = ellipsoid_mesh(r, 40, 80)
    write_obj(os.path.join(d, "ours", "colmap_points.obj"), v - off, f)
    v, f = ellipsoid_mesh(r * 0.9, 16, 32)
    write_obj(os.path.join(d, "ours", "bust_long_tsfm.obj"), v - off, f)
    v, f = ellipsoid_mesh(r * 0.92, 10, 32, cap_only=True)
    write_obj(os.path.join(d, "ours", "scalp_tsfm.obj"), v - off, f)
    return d
