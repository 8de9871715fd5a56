"""Depth-map producer: host-side mirror of Utils/Render_utils.py:render_bust_hair_depth (:310-347) over the CUDA
rasteriser (csrc/sample.cu: mh_render_depth) instead of moderngl / EGL.  Same files: <view>.npy (float32 [H,W,3],
depth * 255, what load_depth reads) and <view>.JPG when capture_imgs, else <view>/bust_hair_depth.png."""
from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .camera import load_cam, parsing_camera
from .pmvo_utils import read_obj

BUST_TO_ORIGIN = np.array([0.006, -1.644, 0.010])          # Render_utils.py:312


class DepthRenderer:
    """The reference's Renderer + BustObj pair for depth frames: meshes are added once, draw(camera) returns the frame."""

    def __init__(self, Height, Width, device="cuda:0"):
        lib()
        self.H, self.W = int(Height), int(Width)
        self.device = torch.device(device)
        self.meshes = []
        self.zbuf = torch.empty((self.H, self.W), dtype=torch.int32, device=self.device)

    def add_mesh(self, vertices, faces):
        v = torch.from_numpy(np.ascontiguousarray(vertices, dtype=np.float32)).to(self.device)
        f = torch.from_numpy(np.ascontiguousarray(faces, dtype=np.int32)).to(self.device)
        self.meshes.append((v, f))

    def draw(self, camera):
        """-> float32 [H,W] device tensor: -z_cam / 2 of the nearest surface, 1 where nothing is drawn (clear colour)."""
        rec = camera.record().numpy()
        out = torch.empty((self.H, self.W), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for k, (v, f) in enumerate(self.meshes or [(None, None)]):
                check(lib().mh_render_depth(stream_ptr(self.device), ptr(v), 0 if v is None else v.size(0), ptr(f),
                                            0 if f is None else f.size(0), rec.ctypes.data, self.H, self.W, ptr(out),
                                            ptr(self.zbuf), 1 if k == 0 else 0), "mh_render_depth")
        return out


def render_bust_hair_depth(colmap_points_path, camera_path, save_root, image_size=[1280, 720], capture_imgs=False,
                           bust_path=None, Headless=True, device="cuda:0"):
    """Render_utils.py:310-347."""
    import cv2
    pts, faces = read_obj(colmap_points_path)
    pts = pts + BUST_TO_ORIGIN
    camera = parsing_camera(load_cam(camera_path))
    R = DepthRenderer(image_size[0], image_size[1], device=device)
    R.add_mesh(pts, faces)
    if bust_path is not None:
        bp, bf = read_obj(bust_path)
        R.add_mesh(bp + BUST_TO_ORIGIN, bf)
    os.makedirs(save_root, exist_ok=True)
    for view, c in camera.items():
        d = R.draw(c)
        depth = d[..., None].expand(-1, -1, 3).cpu().numpy()          # the colour frame: three equal channels
        if capture_imgs:
            depth_save = depth.copy() * 255.
            np.save(os.path.join(save_root, view + '.npy'), depth_save)
            cv2.imwrite(os.path.join(save_root, view + '.JPG'), depth_save)
        else:
            os.makedirs(os.path.join(save_root, view), exist_ok=True)
            cv2.imwrite(os.path.join(save_root, view, 'bust_hair_depth.png'), depth * 255)
