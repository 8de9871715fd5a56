"""Micro-benchmark of mh_voxel_fuse on shell-like points (for ncu and for the HBM roofline of the fusion kernels)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monohair_b200 import synthetic as syn  # noqa: E402
from monohair_b200 import pmvo as P  # noqa: E402
from monohair_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402


def main():
    """no argument: BASELINE configs[1] geometry (256 x 256 x 192, 2.5 mm voxels, 1.25 mm candidate cells);
    `cfg5`: BASELINE configs[4] geometry (512 x 512 x 384, 1.25 mm voxels, 0.625 mm cells: up to 32+ points per voxel)"""
    dev = torch.device("cuda:0")
    big = len(sys.argv) > 1 and sys.argv[1] == "cfg5"
    rng = np.random.default_rng(0)
    if big:
        cand = syn.candidate_points(num_per_grid=4, seed=0, vsize=0.005 / 8, grid=(1024, 1024, 768))
        sel = cand[rng.random(cand.shape[0]) < 0.65]
        P.GRID, P.VOXEL_SIZE = (512, 512, 384), 0.005 / 4
    else:
        cand = syn.candidate_points(num_per_grid=4, seed=0)
        sel = cand[rng.random(cand.shape[0]) < 0.84]
    pts = torch.from_numpy(sel).to(dev).float().contiguous()
    dirs = syn.flow_tangent(pts.double(), syn.RADII).float().contiguous()
    n = pts.size(0)
    gx, gy, gz = P.GRID
    vol = torch.empty((gz, gy, gx, 4), dtype=torch.float32, device=dev)
    _, plane = P.fuse_plane(dev, P.GRID)
    vmin = np.ascontiguousarray(P.VOXEL_MIN)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    nvox = gx * gy * gz
    alg = n * 28 + n * 16 + nvox * 16
    ref = None
    for cap in (0,):
        wsb = lib().mh_voxel_fuse_workspace_bytes(n, gx, gy, gz)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)

        def run():
            check(lib().mh_voxel_fuse(stream_ptr(dev), ptr(pts), ptr(dirs), None, n, vmin.ctypes.data_as(C.c_void_p), float(P.VOXEL_SIZE),
                                      gx, gy, gz, ptr(vol), None, ptr(plane), ptr(ws), wsb), "mh_voxel_fuse")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()                                  # flush L2 (126 MB) between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        occ = int(vol[..., 3].sum().item())
        clean = int(plane[P.FUSE_HDR_BYTES:].view(torch.int64).abs().max().item()) == 0
        hdr = ws[:64].view(torch.int32).cpu().numpy()[[0, 1, 2, 4]]
        same = True
        if ref is None:
            ref = vol.clone()
        else:
            same = bool(torch.equal(ref, vol))
        win, cnt = P.voxel_fuse_winners(pts, dirs, dev, P.GRID, P.VOXEL_MIN, P.VOXEL_SIZE)
        vol2 = P.voxel_scatter(win[: int(cnt.item())], dev, P.GRID)
        same = same and bool(torch.equal(vol2, vol))
        print(f"voxel_fuse: n={n} occupied={occ} crowded-max={hdr[1]} winners={hdr[3]} median {ms*1e3:.1f} us min {min(ts)*1e3:.1f} us  "
              f"algorithmic {alg/1e6:.1f} MB -> {alg/ms/1e6:.0f} GB/s = {alg/ms/1e6/6553:.1%} of 6553  plane_clean={clean} same_volume={same}")
        del ws


if __name__ == "__main__":
    main()
