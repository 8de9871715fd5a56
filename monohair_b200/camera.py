"""Camera model of the hot path: host-side mirror of Utils/Camera_utils.py (same names and conventions).

The per-point projection math itself runs inside the CUDA kernels (csrc/mh_common.cuh); this class only holds
the matrices, exactly as the reference builds them, and packs them into the library's camera record.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


class Camera:
    """Camera_utils.py:10-36.  `pose` is world->camera (the reference passes inv(c2w), :160)."""

    def __init__(self, proj, pose, id, to_tensor=True):
        self.ndc_prj = [float(x) for x in proj]
        self.proj = self.get_projection_matrix(*proj)
        self.pose = pose
        self.id = id
        if to_tensor:
            self.proj = torch.from_numpy(self.proj).type(torch.float)
            self.pose = torch.from_numpy(np.asarray(self.pose)).type(torch.float)

    def get_projection_matrix(self, fx, fy, cx, cy):
        zfar, znear = 100, 0.1
        return np.array([[fx, 0, cx, 0], [0, fy, cy, 0],
                         [0, 0, (-zfar - znear) / (zfar - znear), -2. * zfar * znear / (zfar - znear)],
                         [0, 0, -1, 0]])

    def record(self):
        """32-float camera record (include/monohair_b200.h: MH_CAM_STRIDE).  rinv is torch.linalg.inv of the
        float32 rotation, the same call the reference makes in Camera.reprojection (Camera_utils.py:104)."""
        pose = torch.as_tensor(self.pose, dtype=torch.float).cpu()
        proj = torch.as_tensor(self.proj, dtype=torch.float).cpu()
        rinv = torch.linalg.inv(pose[:3, :3]).contiguous()
        rec = torch.zeros(32, dtype=torch.float)
        rec[0:12] = pose[:3, :].reshape(-1)
        rec[12], rec[13], rec[14], rec[15] = proj[0, 0], proj[1, 1], proj[0, 2], proj[1, 2]
        rec[16:25] = rinv.reshape(-1)
        rec[25:28] = pose[:3, 3]
        return rec


def load_cam(path):
    """Camera_utils.py:141-146."""
    with open(path, 'r') as f:
        cam = json.load(f)
    return cam['cam_list']


def parsing_camera(cam, image_path=None):
    """Camera_utils.py:148-163, including the directory-size sub-sampling (step 1/2/4) and the always-true
    `or c['file']+'.jpg'` of the reference."""
    step = 1
    files = []
    if image_path is not None:
        files = os.listdir(image_path)
        if len(files) > 500:
            step = 4
        elif len(files) > 300:
            step = 2
    camera = {}
    for c in cam[::step]:
        camera[c['file']] = Camera(c['ndc_prj'], np.linalg.inv(np.array(c['pose'])), c['file'])
    return camera


def cameras_from_scene(scene):
    return {c["file"]: Camera(c["ndc_prj"], np.linalg.inv(np.array(c["pose"])), c["file"]) for c in scene.cams}
