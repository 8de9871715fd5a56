"""The other BASELINE.json configs, measured beside the headline PMVO job: configs[0] (Gabor bank on a 512x512 frame, the
reference's CPU-runnable case; plus the float32 pipeline bank at 512x512 / 1080p) and configs[3] (HairGrow, 100 k strands
through a 256x256x192 volume, .mat -> .hair), each next to the CPU oracle port timed on a bounded sample.
bench.py calls gabor_workload() / hairgrow_workload() at N=1 and puts the dicts on its JSON line under
"other_workloads"; `python tools/bench_extra.py [gabor|hairgrow]` prints them alone."""
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


def gabor_workload(cpu=True, sizes=((512, 512), (1080, 1920))):
    """-> dict of measurements (BASELINE configs[0] + the float32 pipeline bank)."""
    from monohair_b200 import gabor as MG
    from oracle import gabor_oracle as G
    out = {}
    img = texture_rgb(512, 512)
    ks = MG.generate_gabor_filters(1.8, 2.4, 0.23, 180)
    f = lambda: MG.calc_orients(img.astype(np.float64), ks)
    ms64 = ev_time(f, reps=3, warm=1)
    r = {"config": "calc_orientation_maps.calc_orients, 180 filters, 512x512 frame, float64 (BASELINE configs[0])",
         "ms_per_frame": ms64, "frames_per_s": 1e3 / ms64}
    if cpu:
        t = time.time(); F_o = G.calc_orients(img.astype(np.float64), ks); cpu64 = time.time() - t
        F = f().cpu().numpy()
        r.update(cpu_numpy_s_per_frame=cpu64, cores=os.cpu_count(), max_abs_diff_vs_cpu=float(np.abs(F - F_o).max()),
                 argmax_identical=float(np.mean(F.argmax(0) == F_o.argmax(0))))
    out["calc_orientation_maps_512x512_f64"] = r
    m = MG.calOrientationGabor()
    for (H, W) in sizes:
        x = torch.rand((1, 1, H, W), device="cuda") * 0.2 - 0.1
        ms = ev_time(lambda: m(x), reps=5, warm=2)
        flops = 2.0 * 289 * 180 * H * W
        out[f"calOrientationGabor_{H}x{W}_f32"] = {"ms_per_frame": ms, "fp32_tflops": flops / ms / 1e9, "frames_per_s": 1e3 / ms,
                                                   "algorithmic_flops_per_pixel": 2 * 289 * 180}
    if cpu:
        x = torch.rand((1, 1, 512, 512)) * 0.2 - 0.1
        t = time.time(); G.gabor_orientation(x[0, 0].numpy()); c = time.time() - t
        out["calOrientationGabor_512x512_f32"]["cpu_oracle_port_s_per_frame"] = c
        out["calOrientationGabor_512x512_f32"]["cores"] = os.cpu_count()
    return out


def bench_gabor():
    for k, v in gabor_workload(sizes=((512, 512), (1080, 1920), (2160, 3840))).items():
        print(json.dumps({k: v}))


def hairgrow_workload(cpu=True, shell_mm=6.0, tmpdir=None, n_roots=90000):
    """BASELINE configs[3]: strands through a 256x256x192 orientation field; the shell is thick enough for >= 100 k
    accepted strands.  Also times the stage end to end from the .mat pair to scalp_segment.hair."""
    import tempfile
    from monohair_b200 import pmvo as P
    from monohair_b200.hairgrow import HairGrowing, save_hair_strands
    from oracle import hairgrow_oracle as H
    occ, ori = syn.orientation_volume(device="cuda:0", shell_mm=shell_mm)
    vol = torch.zeros((192, 256, 256, 4), device="cuda:0")
    o = torch.from_numpy(ori).cuda().float()
    vol[..., 0] = o[..., 0].permute(2, 1, 0)
    vol[..., 1] = -o[..., 1].permute(2, 1, 0)
    vol[..., 2] = -o[..., 2].permute(2, 1, 0)
    vol[..., 3] = torch.from_numpy(occ).cuda().float().permute(2, 1, 0)
    del o
    hg = HairGrowing(volume=vol, device="cuda:0")
    # scalp roots on a smaller ellipsoid, voxel coordinates
    rng = np.random.default_rng(0)
    d = rng.normal(size=(400000, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True); d = d[d[:, 1] > 0.2][:n_roots]
    r = np.array(syn.RADII) * 0.9
    p = d * r
    nrm = p / (r * r); nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    flip = np.array([1.0, -1.0, -1.0])
    roots = torch.from_numpy(((p * flip - syn.BBOX_MIN) / syn.COARSE_VSIZE).astype(np.float32)).cuda()
    normals = torch.from_numpy((nrm * flip).astype(np.float32)).cuda()
    torch.manual_seed(0)
    hg.GenerateGuideStrandFromScalp(roots[:2000], normals[:2000], None, 0.85)        # warm-up (allocator, kernels)
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
    pts_a, off_a, ln_a = hg._trace_batch(seeds, 0.85)
    steps = int(ln_a.sum().item())
    flag0 = torch.zeros((hg.gz, hg.gy, hg.gx), dtype=torch.float32, device="cuda:0")
    rp, ro, rl = hg._scalp_batch(roots, normals, 0.85)
    hg._accept(rp, ro, rl, None, flag0, 1)
    ms_accept = ev_time(lambda: hg._accept(pts_a, off_a, ln_a, seeds, flag0.clone(), 0), reps=3, warm=1)
    out = {"config": "HairGrow GenerateGuideStrandFromScalp, 256x256x192 volume (BASELINE configs[3])",
           "occupied_voxels": M, "seeds": int(n_roots + 2 * M), "strands": len(strands), "num_root": num_root, "points": n_pts,
           "s_total": dt, "strands_per_s": len(strands) / dt, "trace_only_ms_per_pass": ms_trace,
           "trace_seeds_per_s": M / (ms_trace * 1e-3), "trace_steps_per_pass": steps,
           # one dependent 16 B voxel fetch + 12 B point write per step: a latency / random-sector kernel, not a streaming one
           "trace_GBps_algorithmic": steps * 28 / (ms_trace * 1e-3) / 1e9, "accept_ms_per_pass": ms_accept}
    # .mat pair -> scalp_segment.hair, the files the reference's stage reads and writes
    import scipy.io
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        occ_m, ori_m = P.volume_to_mat(vol)
        scipy.io.savemat(td + "/Ori3D.mat", {"Ori": ori_m.cpu().numpy()})
        scipy.io.savemat(td + "/Occ3D.mat", {"Occ": occ_m.cpu().numpy()})
        del occ_m, ori_m
        torch.manual_seed(0)
        torch.cuda.synchronize()
        t = time.time()
        hg2 = HairGrowing(td + "/Occ3D.mat", td + "/Ori3D.mat", device="cuda:0")
        t_load = time.time() - t
        st2, _ = hg2.GenerateGuideStrandFromScalp(roots, normals, None, 0.85)
        world = hg2.VoxelToWorld(st2, np.array([0.006, -1.644, 0.010]))
        save_hair_strands(td + "/scalp_segment.hair", world)
        torch.cuda.synchronize()
        out.update(mat_to_hair_s=time.time() - t, mat_load_s=t_load, hair_file_MB=os.path.getsize(td + "/scalp_segment.hair") / 1e6)
    if cpu:
        volc = H.Volume(vol[..., 3].cpu().numpy(), vol[..., :3].permute(3, 0, 1, 2).contiguous().cpu().numpy())
        sd = seeds.cpu().numpy()[:: max(1, M // 300)][:300].copy()
        flag = np.zeros_like(volc.occ)
        t = time.time(); kept = 0
        for i in range(sd.shape[0]):
            s = H.trace(volc, sd[i], flag, 0.85, np.zeros(3, np.float32)); kept += s is not None
        c = time.time() - t
        out.update(cpu_oracle_seeds_per_s=sd.shape[0] / c, cpu_sample=f"{sd.shape[0]} seeds, 1 core (scalar port)")
    return out


def bench_hairgrow():
    print(json.dumps(hairgrow_workload()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gabor", "all"):
        bench_gabor()
    if what in ("hairgrow", "all"):
        bench_hairgrow()
