"""Secondary benchmarks (not the driver's bench.py contract): BASELINE.json configs[0] (Gabor bank on a 512x512 frame,
the reference's CPU-runnable case) and configs[3] (HairGrow through a 256x256x192 volume), each next to the CPU oracle
port timed on a bounded sample.  Prints one JSON line per config."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monohair_b200 import synthetic as syn  # noqa: E402


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def texture_rgb(H, W, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W))
    for _ in range(200):
        th, wl, ph = rng.uniform(0, np.pi), rng.uniform(3, 6), rng.uniform(0, 2 * np.pi)
        cx, cy, s = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(10, 40)
        img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s)) * np.cos(2 * np.pi * (xx * np.cos(th) + yy * np.sin(th)) / wl + ph)
    img = np.clip(128 + 40 * img + rng.normal(0, 4, (H, W)), 0, 255)
    return np.repeat(img[..., None], 3, -1).astype(np.uint8)


def bench_gabor():
    from monohair_b200 import gabor as MG
    from oracle import gabor_oracle as G
    img = texture_rgb(512, 512)
    ks = MG.generate_gabor_filters(1.8, 2.4, 0.23, 180)
    f = lambda: MG.calc_orients(img.astype(np.float64), ks)
    ms64 = ev_time(f, reps=3, warm=1)
    t = time.time(); F_o = G.calc_orients(img.astype(np.float64), ks); cpu64 = time.time() - t
    F = f().cpu().numpy()
    print(json.dumps({"config": "calc_orientation_maps Gabor bank, 512x512 frame, float64 (BASELINE configs[0])",
                      "b200_ms_per_frame": ms64, "cpu_numpy_s_per_frame": cpu64, "cores": os.cpu_count(),
                      "max_abs_diff_vs_cpu": float(np.abs(F - F_o).max()), "argmax_identical": float(np.mean(F.argmax(0) == F_o.argmax(0)))}))
    m = MG.calOrientationGabor()
    for (H, W) in ((512, 512), (1080, 1920), (2160, 3840)):
        x = torch.rand((1, 1, H, W), device="cuda") * 0.2 - 0.1
        ms = ev_time(lambda: m(x), reps=5, warm=2)
        flops = 2.0 * 289 * 180 * H * W
        print(json.dumps({"config": f"calOrientationGabor.forward {H}x{W} float32", "b200_ms_per_frame": ms,
                          "fp32_tflops": flops / ms / 1e9, "frames_per_s": 1e3 / ms}))
    x = torch.rand((1, 1, 512, 512)) * 0.2 - 0.1
    t = time.time(); G.gabor_orientation(x[0, 0].numpy()); cpu = time.time() - t
    print(json.dumps({"config": "calOrientationGabor.forward 512x512, CPU oracle port", "cpu_s_per_frame": cpu, "cores": os.cpu_count()}))


def bench_hairgrow():
    from monohair_b200.hairgrow import HairGrowing
    from oracle import hairgrow_oracle as H
    occ, ori = syn.orientation_volume(device="cuda:0")
    vol = torch.zeros((192, 256, 256, 4), device="cuda:0")
    o = torch.from_numpy(ori).cuda().float()
    vol[..., 0] = o[..., 0].permute(2, 1, 0)
    vol[..., 1] = -o[..., 1].permute(2, 1, 0)
    vol[..., 2] = -o[..., 2].permute(2, 1, 0)
    vol[..., 3] = torch.from_numpy(occ).cuda().float().permute(2, 1, 0)
    hg = HairGrowing(volume=vol, device="cuda:0")
    # scalp roots on a smaller ellipsoid, voxel coordinates
    rng = np.random.default_rng(0)
    d = rng.normal(size=(200000, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True); d = d[d[:, 1] > 0.2][:60000]
    r = np.array(syn.RADII) * 0.9
    p = d * r
    nrm = p / (r * r); nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    flip = np.array([1.0, -1.0, -1.0])
    roots = torch.from_numpy(((p * flip - syn.BBOX_MIN) / syn.COARSE_VSIZE).astype(np.float32)).cuda()
    normals = torch.from_numpy((nrm * flip).astype(np.float32)).cuda()
    torch.manual_seed(0)
    torch.cuda.synchronize()
    t = time.time()
    strands, num_root = hg.GenerateGuideStrandFromScalp(roots, normals, None, 0.85)
    torch.cuda.synchronize()
    dt = time.time() - t
    n_pts = int(sum(s.shape[0] for s in strands))
    M = int((vol[..., 3] > 0).sum().item())
    # trace-only timing (count + write passes over all occupied voxels)
    seeds = hg._positive_seeds() + 0.6
    ms_trace = ev_time(lambda: hg._trace_batch(seeds, 0.85), reps=3, warm=1)
    # ordered acceptance of one pass over all occupied voxels (flag volume as the scalp pass leaves it)
    pts_a, off_a, ln_a = hg._trace_batch(seeds, 0.85)
    flag0 = torch.zeros((hg.gz, hg.gy, hg.gx), dtype=torch.float32, device="cuda:0")
    rp, ro, rl = hg._scalp_batch(roots, normals, 0.85)
    hg._accept(rp, ro, rl, None, flag0, 1)
    ms_accept = ev_time(lambda: hg._accept(pts_a, off_a, ln_a, seeds, flag0.clone(), 0), reps=3, warm=1)
    # CPU oracle on a bounded sample
    volc = H.Volume(vol[..., 3].cpu().numpy(), vol[..., :3].permute(3, 0, 1, 2).contiguous().cpu().numpy())
    sd = seeds.cpu().numpy()[:: max(1, M // 300)][:300].copy()
    flag = np.zeros_like(volc.occ)
    t = time.time(); kept = 0
    for i in range(sd.shape[0]):
        s = H.trace(volc, sd[i], flag, 0.85, np.zeros(3, np.float32)); kept += s is not None
    cpu = time.time() - t
    print(json.dumps({"config": "HairGrow GenerateGuideStrandFromScalp, 256x256x192 volume (BASELINE configs[3])",
                      "occupied_voxels": M, "seeds": int(60000 + 2 * M), "strands": len(strands), "num_root": num_root,
                      "points": n_pts, "b200_s_total": dt, "strands_per_s": len(strands) / dt,
                      "trace_only_ms_per_pass": ms_trace, "trace_seeds_per_s": M / (ms_trace * 1e-3),
                      "accept_ms_per_pass": ms_accept,
                      "cpu_oracle_seeds_per_s": sd.shape[0] / cpu, "cpu_sample": f"{sd.shape[0]} seeds, 1 core (scalar port)"}))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gabor", "all"):
        bench_gabor()
    if what in ("hairgrow", "all"):
        bench_hairgrow()
