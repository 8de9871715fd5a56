#!/usr/bin/env python
"""bench.py -- PMVO points ("voxels")/s on the big_wavy1-like synthetic capture (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale small|full]

One "step" = one whole PMVO job on the scene: filter_points over all candidates -> forward (patch-based
multi-view optimisation) over the surface points -> kNN-medoid refine + re-scoring -> orientation of the
near-surface points -> voxel fusion into the 256x256x192 occupancy/orientation volume.
`value` = points that went through the optimisation / step time, maps resident in HBM.
`e2e`   = same job through the reference-facing API from (pinned) HOST buffers: H2D of every view's maps and of
          the candidates, and D2H of the fused volume + per-point results, inside the timed region.
Rank 0 prints ONE JSON line.  `--impl reference` times the CPU oracle port of the reference (oracle/) on a bounded
sample of the same workload (the reference itself is Python and cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: PMVO configs/reconstruct/big_wavy1 (image_size [1920,1080], patch 7,
    # conf_threshold .15, threshold .025, num_sample_per_grid 4), 60 views, grid 256x256x192
    "full": dict(name="pmvo_big_wavy1_synth_60v_1920x1080", V=60, H=1920, W=1080, patch=7, conf_thr=0.15, thr=0.025,
                 visible_thr=1, num_per_grid=4, n_cells=None),
    # BASELINE.json configs[4]: PMVO + HairGrow, 120 views @ 2160p, 512 x 512 x 384 grid (0.625 mm candidate cells, 1.25 mm
    # voxels), meant for 8 GPUs (32 GB of view maps per rank); run by hand: --scale cfg5 --no-e2e --no-cpu --no-extra
    "cfg5": dict(name="pmvo_hairgrow_120v_2160x3840_grid512x512x384", V=120, H=2160, W=3840, patch=7, conf_thr=0.15, thr=0.025,
                 visible_thr=1, num_per_grid=4, n_cells=None, fine_vsize=0.005 / 8, fine_grid=(1024, 1024, 768),
                 grid=(512, 512, 384), voxel_size=0.005 / 4, hairgrow=True),
    "small": dict(name="pmvo_small_24v_480x270", V=24, H=480, W=270, patch=7, conf_thr=0.15, thr=0.025,
                  visible_thr=1, num_per_grid=2, n_cells=20000),
}
CPU_SAMPLE_CANDIDATES = 4600         # cpu_baseline leg of the B200 arm: ONE step, ~2000 optimised points (~25 s of CPU)
REF_ARM_CPU_SECONDS = 150.0          # reference arm: all --steps together stay near this much CPU time
REF_ARM_POINTS_PER_S = 35.0          # planning figure for the oracle port (candidates/s incl. filter), 16 cores


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(cfg, device, seed=0, views=None):
    from monohair_b200 import synthetic as syn
    t = time.time()
    sc = syn.make_scene(V=cfg["V"], H=cfg["H"], W=cfg["W"], seed=seed, device=device, views=views)
    kw = {}
    if "fine_vsize" in cfg:
        kw = dict(vsize=cfg["fine_vsize"], grid=cfg["fine_grid"])
    cand = syn.candidate_points(n_cells=cfg["n_cells"], num_per_grid=cfg["num_per_grid"], seed=seed, **kw)
    scalp = syn.scalp_vertices(2000, seed=seed)
    log(f"workload {cfg['name']}: scene+candidates in {time.time() - t:.1f}s, {cand.shape[0]} candidate points")
    return sc, cand, scalp


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_step(vm, cand_sample, cfg, scalp):
    """One bounded step of the CPU oracle port: the same stages on `cand_sample`.  Returns (#points optimised, s)."""
    from scipy.spatial import KDTree
    from oracle import pmvo_oracle as O
    t0 = time.time()
    tree, smax = KDTree(data=scalp), scalp.max(0)
    pts = torch.from_numpy(cand_sample).float()
    s, f, _ = O.filter_points(vm, pts, cfg["patch"], cfg["visible_thr"], cfg["conf_thr"])
    sp = cand_sample[s.numpy()]
    p, o, l, hc = O.forward(vm, sp, cfg["patch"], cfg["conf_thr"])
    k = min(100, len(sp))
    p2, o2, l2 = O.refine_points(vm, p.numpy(), o.numpy(), l.numpy(), cfg["patch"], cfg["visible_thr"], cfg["conf_thr"],
                                 tree, smax, k=k)
    idx = np.where(l2 < cfg["thr"])[0]
    fu = cand_sample[f.numpy()]
    if len(fu) and len(idx) >= k:
        fp, fo = O.unvisible_orientation(vm, p2[idx], o2[idx], fu, cfg["visible_thr"], tree, smax, k=k)
    else:
        fp, fo = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
    O.voxel_fuse(np.concatenate([p2[idx], fp]), np.concatenate([o2[idx], fo]))
    return len(sp), time.time() - t0


def config_dict(cfg):
    """the workload, in the same words for both arms (the driver compares the two `config` objects)."""
    return {"workload": cfg["name"], "views": cfg["V"], "image": [cfg["H"], cfg["W"]], "patch": cfg["patch"],
            "grid": list(cfg.get("grid", (256, 256, 192))), "num_sample_per_grid": cfg["num_per_grid"],
            "cache": "inputs larger than L2 (%.1f GB of view maps, gathered)" % ((cfg["V"] * cfg["H"] * cfg["W"] * 32) / 1e9)}


def cpu_sample(cand, n):
    """every (N/n)-th candidate of the covered range: same spatial distribution as the full job."""
    stride = max(1, cand.shape[0] // n)
    return np.ascontiguousarray(cand[::stride][:n])


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pmvo_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda" if torch.cuda.is_available() else "cpu"          # only for generating the synthetic maps
    sc, cand, scalp = make_workload(cfg, dev)
    vm = O.ViewMaps.from_scene(sc)
    # bounded sample per step, sized so that the whole --steps run ends within a few minutes
    per_step = int(min(CPU_SAMPLE_CANDIDATES, max(320, REF_ARM_CPU_SECONDS * REF_ARM_POINTS_PER_S * 4 / max(args.steps, 1))))
    sample = cpu_sample(cand, per_step)
    for _ in range(min(args.warmup, 2)):
        cpu_port_step(vm, sample[:64], cfg, scalp)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, t = cpu_port_step(vm, sample, cfg, scalp)
        n_tot += n
        t_tot += t
    val = n_tot / t_tot
    line = {"impl": "reference", "metric": "pmvo_points_per_s", "value": val, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg),
            "cpu_baseline": {"value": val, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(sample)} candidate points/step ({n_tot // max(args.steps,1)} optimised), "
                                       f"all {cfg['V']} views at full resolution; torch-CPU oracle port of the reference"},
            "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, cfg):
    import torch.distributed as dist
    from monohair_b200 import _lib, pipeline
    from monohair_b200 import pmvo as P
    from monohair_b200.camera import cameras_from_scene
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    views = None
    if world > 1 and cfg.get("hairgrow"):                      # big captures: a rank generates only the views it uploads
        per = (cfg["V"] + world - 1) // world
        views = range(min(rank * per, cfg["V"]), min((rank + 1) * per, cfg["V"]))
    sc, cand_np, scalp = make_workload(cfg, dev, views=views)
    grid = tuple(cfg.get("grid", (256, 256, 192)))
    voxel_size = float(cfg.get("voxel_size", 0.005 / 2))
    P.scalp_tree, P.scalp_max = scalp, scalp.max(0)
    cams = cameras_from_scene(sc)
    kw = dict(device=dev, image_size=[cfg["H"], cfg["W"]], patch_size=cfg["patch"], visible_threshold=cfg["visible_thr"],
              conf_threshold=cfg["conf_thr"])
    pm = P.PMVO.from_u8(cams, sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, **kw)
    cand64 = torch.from_numpy(cand_np).to(dev).contiguous()        # as loaded (float64): queries of the near-surface kNN
    cand = cand64.float().contiguous()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------
    stage_ms = {}
    stats = {}

    all_events = []

    def one_step(record):
        evs = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(dev))
            evs.append((name, e))
        out = pipeline.pmvo_job_device(pm, cand, cfg["thr"], stats=stats, mark=mark, grid=grid, voxel_size=voxel_size,
                                       cand64=cand64 if cand64.dtype == torch.float64 else None)
        if record:
            all_events.append(evs)          # elapsed times are read after the timed region (no extra sync inside it)
        return out

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = L.mh_launch_count()
    barrier()
    profiling = os.environ.get("MH_PROFILE") == "1"       # ncu --profile-from-start off: only the timed steps
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    t_wall0 = time.time()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(torch.cuda.current_stream(dev))
    for _ in range(args.steps):
        out = one_step(True)
    e_end.record(torch.cuda.current_stream(dev))
    barrier()
    if profiling:
        torch.cuda.cudart().cudaProfilerStop()
    t_wall1 = time.time()
    launches = L.mh_launch_count() - launches0
    ms_total = e_start.elapsed_time(e_end)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    for evs in all_events:
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            stage_ms.setdefault(n1, []).append(e0.elapsed_time(e1))
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    n_opt = out["n_optimized"]
    # bit-level checksums of the results (int32 view, summed in int64): equal across N means the sharded job reproduces
    # the single-GPU one bit for bit
    cks = lambda t: int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())
    checksums = {"volume": cks(out["volume"]), "select_o": cks(out["select_o"]), "min_loss": cks(out["min_loss"]),
                 "refine_o": cks(out["refine_o"]), "occupied": int(out["volume"][..., 3].sum().item())}
    value = n_opt / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (optimize) and of the HBM-bound ones --------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if peaks else "fallback"
    med = {k: float(np.median(v)) for k, v in stage_ms.items()}
    n_opt_local = (n_opt + world - 1) // world
    PP = cfg["patch"] ** 2
    opt_bytes = n_opt_local * cfg["V"] * (PP * 12 + 4 + 4 + 8 + 4)              # 608 B per (point, view) at P=7
    n_cov = stats["n_candidates"] // 30 * (30 if stats["n_candidates"] % 30 == 0 else 31)
    n_cov = min(n_cov, stats["n_candidates"])
    # filter_points: what this implementation has to read per (point, view) = 8 B {depth, mask'} + 4 B (PxP max conf);
    # + 12 B point + 20 B counters per point.  (The reference's own formulation gathers 204 B per pair)
    filt_bytes = ((n_cov + world - 1) // world) * (cfg["V"] * 12 + 12 + 20)
    n_fused = stats["n_selected"] + stats["n_fu"]
    nvox = grid[0] * grid[1] * grid[2]
    fuse_bytes = n_fused * 28 + n_fused * 2 * 8 + nvox * 16                      # SURVEY §8d voxel fusion
    # optimize_kernel (83 % of the step) is instruction-issue bound, not HBM bound (profiles/r2_optimize_ncu.md):
    # its ruler is the SM's measured issue rate (tools/microbench_peaks.cu -> profiles/r2_fp32_peaks.json), against the
    # warp-instructions ncu counted for this workload (per optimised point; profiles/r2_traffic.json).  The HBM figure
    # stays as a note, and the HBM-bound kernels of the path are listed beside it.
    traffic, instr_per_point, issue_peak, fp32_peak = None, None, None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["optimize_kernel"]
        if tr["workload"] == cfg["name"]:
            traffic = tr["dram_bytes_per_launch"] * (n_opt_local / tr["points_per_launch"])
            instr_per_point = tr["warp_instructions_per_launch"] / tr["points_per_launch"]
        pk = json.load(open(os.path.join(ROOT, "profiles", "r2_fp32_peaks.json")))
        issue_peak = max(v["g_warp_instr_per_s"] for k, v in pk.items() if isinstance(v, dict))
        fp32_peak = pk["ffma2"]["tflops"]
    except Exception:
        pass
    opt_s = med["optimize"] * 1e-3
    if instr_per_point and issue_peak:
        ach = instr_per_point * n_opt_local / opt_s / 1e9
        roofline = {"bound": "fp32_issue", "kernel": "optimize_kernel (PMVO.forward)", "achieved": ach, "peak": issue_peak,
                    "unit": "G warp-instr/s", "frac": ach / issue_peak, "traffic": traffic,
                    "peak_source": "measured (tools/microbench_peaks.cu on this pool's B200, profiles/r2_fp32_peaks.json; "
                                   "FP32 FMA peak %.1f TFLOP/s)" % fp32_peak,
                    "ms": med["optimize"], "warp_instructions_per_point": instr_per_point,
                    "hbm_note": {"algorithmic_GBps": opt_bytes / opt_s / 1e9, "frac_of_hbm_peak": opt_bytes / opt_s / 1e9 / hbm_peak,
                                 "hbm_peak_GBps": hbm_peak, "why": "608 B per (point, view) in the reference's fp32 terms are gathered "
                                 "once and reused ~1e5 instructions long; DRAM is 0.08x of that"}}
    else:
        roofline = {"bound": "hbm", "kernel": "optimize_kernel (PMVO.forward)", "achieved": opt_bytes / opt_s / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": opt_bytes / opt_s / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "ms": med["optimize"]}
    roofline["hbm_bound_kernels"] = {
        "voxel_fuse": {"achieved": fuse_bytes / (med["fuse"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": fuse_bytes / (med["fuse"] * 1e-3) / 1e9 / hbm_peak, "ms": med["fuse"], "peak_source": peak_src},
        "filter_count": {"achieved": filt_bytes / (med["filter"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": filt_bytes / (med["filter"] * 1e-3) / 1e9 / hbm_peak, "ms": med["filter"], "peak_source": peak_src}}

    # ---- BASELINE configs[4]: the strand stage on the fused volume, straight from device memory (rank 0) ------------
    hairgrow = None
    if cfg.get("hairgrow") and rank == 0:
        import contextlib
        from monohair_b200 import synthetic as syn
        from monohair_b200.hairgrow import HairGrowing
        with contextlib.redirect_stdout(sys.stderr):
            hg = HairGrowing(volume=out["volume"], device=dev)
            rng = np.random.default_rng(0)
            d = rng.normal(size=(400000, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True); d = d[d[:, 1] > 0.2][:60000]
            r = np.array(syn.RADII) * 0.9
            pp = d * r
            nrm = pp / (r * r); nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
            flip = np.array([1.0, -1.0, -1.0])
            roots = torch.from_numpy(((pp * flip - syn.BBOX_MIN) / voxel_size).astype(np.float32)).to(dev)
            normals = torch.from_numpy((nrm * flip).astype(np.float32)).to(dev)
            torch.manual_seed(0)
            torch.cuda.synchronize()
            t0 = time.time()
            strands, num_root = hg.GenerateGuideStrandFromScalp(roots, normals, None, 0.85)
            torch.cuda.synchronize()
            hairgrow = {"s": time.time() - t0, "strands": len(strands), "num_root": int(num_root),
                        "points": int(sum(x.shape[0] for x in strands)), "occupied_voxels": int(out["volume"][..., 3].sum().item())}
            del strands, hg
    if world > 1 and cfg.get("hairgrow"):
        dist.barrier()

    # ---- end to end from host buffers ----------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Ori, Conf = sc.ref_ori_conf()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_depth = {k: pin(v) for k, v in sc.ref_depths().items()}
        h_ori = {k: pin(v) for k, v in Ori.items()}
        h_conf = {k: pin(v) for k, v in Conf.items()}
        h_mask = {k: pin(v) for k, v in sc.ref_masks().items()}
        h_cand = pin(cand_np)
        del Ori, Conf
        h2d = sum(t.numel() * t.element_size() for d in (h_depth, h_ori, h_conf, h_mask) for t in d.values()) \
            + h_cand.numel() * h_cand.element_size()
        del pm
        torch.cuda.empty_cache()
        kw2 = dict(image_size=[cfg["H"], cfg["W"]], patch_size=cfg["patch"], visible_threshold=cfg["visible_thr"],
                   conf_threshold=cfg["conf_thr"], threshold=cfg["thr"], device=dev)
        n_e2e = max(1, args.steps)
        host = pipeline.pmvo_job_host(cams, h_depth, h_ori, h_conf, h_mask, h_cand, **kw2)      # warm-up
        barrier()
        t0 = time.time()
        for _ in range(n_e2e):
            host = pipeline.pmvo_job_host(cams, h_depth, h_ori, h_conf, h_mask, h_cand, **kw2)
        barrier()
        t_e2e = (time.time() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([t_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        d2h = sum(host[k].numel() * host[k].element_size() for k in ("volume", "select_o", "min_loss", "high_conf") if k in host)   # rank 0 reads back
        e2e = {"value": host["n_optimized"] / t_e2e, "unit": "points/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "s_per_step": t_e2e,
               "api": "PMVO(camera,depths,Ori,Conf,masks) with the reference loaders' float64/float32 arrays (pinned) + "
                      "filter/forward/refine/fuse; volume and per-point results read back"}
        del h_depth, h_ori, h_conf, h_mask
        # the same job fed with the maps in their FILE formats (8-bit orientation / confidence / mask images, float32 depth
        # channel; PMVO.from_u8 decodes them inside the pack kernel): what a pipeline that reads the
        # stage files directly pays.  Reported beside `e2e`, which keeps the reference loaders' float64 arrays.
        if torch.is_tensor(sc.depth):
            u_depth, u_ori, u_conf, u_mask = [t.cpu().contiguous().pin_memory() for t in (sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8)]
        else:
            u_depth, u_ori, u_conf, u_mask = [pin(a) for a in (sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8)]
        h2d_u8 = sum(t.numel() * t.element_size() for t in (u_depth, u_ori, u_conf, u_mask)) + h_cand.numel() * h_cand.element_size()
        torch.cuda.empty_cache()
        host = pipeline.pmvo_job_host(cams, u_depth, u_ori, u_conf, u_mask, h_cand, u8=True, **kw2)                 # warm-up
        barrier()
        t0 = time.time()
        for _ in range(n_e2e):
            host = pipeline.pmvo_job_host(cams, u_depth, u_ori, u_conf, u_mask, h_cand, u8=True, **kw2)
        barrier()
        t_u8 = (time.time() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([t_u8], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_u8 = float(t.item())
        e2e["file_format_maps"] = {"value": host["n_optimized"] / t_u8, "unit": "points/s", "h2d_bytes_per_step": int(h2d_u8),
                                   "s_per_step": t_u8, "api": "PMVO.from_u8(camera, depth f32, best_ori u8, conf u8, mask u8)"}
        del u_depth, u_ori, u_conf, u_mask

    # ---- CPU baseline (rank 0, N=1) --------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import pmvo_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        vm = O.ViewMaps.from_scene(sc)
        sample = cpu_sample(cand_np, CPU_SAMPLE_CANDIDATES)
        n, t = cpu_port_step(vm, sample, cfg, scalp)
        cpu = {"value": n / t, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{len(sample)} candidate points ({n} optimised) in {t:.1f}s, all {cfg['V']} views; "
                         f"torch-CPU oracle port of the reference"}

    # ---- the other BASELINE configs (rank 0, N=1, outside the timed region) -------------------------------------
    other = None
    if rank == 0 and world == 1 and not args.no_extra:
        import contextlib
        try:
          with contextlib.redirect_stdout(sys.stderr):            # stdout carries the one JSON line only
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_extra
            torch.cuda.empty_cache()
            other = dict(bench_extra.gabor_workload(cpu=not args.no_cpu))
            other["hairgrow_256x256x192"] = bench_extra.hairgrow_workload(cpu=not args.no_cpu)
            # the fusion on the 256^3 grid of BASELINE configs[2]/[3] (same points)
            g3, vmin3 = (256, 256, 256), (-0.32, -0.32, -0.32)
            sel = out["refine_loss"] < cfg["thr"]
            ap = torch.cat([out["select_p"][sel], out["fu_points"]], 0).contiguous()
            ao = torch.cat([out["refine_o"][sel], out["fu_ori"]], 0).contiguous()
            ms3 = bench_extra.ev_time(lambda: P.voxel_fuse(ap, ao, dev, g3, vmin3), reps=5, warm=2)
            b3 = ap.size(0) * 44 + 256 ** 3 * 16
            other["voxel_fuse_256cubed"] = {"ms": ms3, "algorithmic_MB": b3 / 1e6, "GBps": b3 / ms3 / 1e6, "frac_of_hbm_peak": b3 / ms3 / 1e6 / hbm_peak}
        except Exception as e:  # pragma: no cover
            other = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "pmvo_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(cfg),
                "run": {"candidates": stats["n_candidates"], "optimised": n_opt, "selected": stats["n_selected"],
                        "near_surface": stats["n_fu"],
                        "parallelism": "1 GPU" if world == 1 else
                        f"points sharded over {world} GPUs (filter, forward, kNN, re-score, near-surface medoids); "
                        f"chunk-ordered medoid sweep {pipeline._sweep_mode(dist, dev)} "
                        f"({'one kernel per rank storing finished points into every peer copy over NVLink' if pipeline._sweep_mode(dist, dev) == 'peer' and not pipeline._SYMM_BROKEN else 'every rank sweeps all points'}); "
                        f"fusion {os.environ.get('MH_FUSE_DIST', 'replicated')}"},
                "checksums": checksums,
                "stage_ms": med, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "other_workloads": other, "hairgrow_on_fused_volume": hairgrow}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", default=os.environ.get("MH_BENCH_SCALE", "full"), choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--override", default=None, help='JSON dict merged into the workload (smoke tests of big configs, e.g. {"V": 24, "H": 540, "W": 960})')
    ap.add_argument("--no-extra", action="store_true", help="skip the Gabor / HairGrow / 256^3 workloads reported beside the PMVO job")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, int(os.environ.get("MH_BENCH_MIN_WARMUP", "3")))   # profiling runs may lower it
    cfg = dict(WORKLOADS[args.scale])
    if args.override:
        cfg.update(json.loads(args.override))
        cfg["name"] += "_override"
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
