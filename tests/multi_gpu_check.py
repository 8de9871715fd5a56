"""Run under torchrun on >= 2 GPUs (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
Checks that the sharded PMVO job (points sharded over the ranks; fusion replicated or, with MH_FUSE_DIST=winners, sharded
by voxel slab with an all-gather of the per-voxel winners over NCCL) reproduces the single-GPU job bit-for-bit: every
stage is per-point / per-voxel independent, so sharding must not change a single value.  Run by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monohair_b200 import pipeline, synthetic as syn  # noqa: E402
from monohair_b200 import pmvo as P  # noqa: E402
from monohair_b200.camera import cameras_from_scene  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    sc = syn.make_scene(V=24, H=270, W=480, seed=5)
    cand = syn.candidate_points(n_cells=6000, num_per_grid=2, seed=5)
    scalp = syn.scalp_vertices(500, seed=5)
    P.scalp_tree, P.scalp_max = scalp, scalp.max(0)
    pm = P.PMVO.from_u8(cameras_from_scene(sc), sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, device=dev,
                        image_size=[sc.H, sc.W], patch_size=7, visible_threshold=1, conf_threshold=0.15)
    # view-sharded upload (+ all-gather of the packed planes over NCCL) == every rank packing every view
    P.SHARD_VIEW_UPLOAD = False
    Ori, Conf = sc.ref_ori_conf()
    pm_full = P.PMVO(cameras_from_scene(sc), sc.ref_depths(), Ori, Conf, sc.ref_masks(), device=dev,
                     image_size=[sc.H, sc.W], patch_size=7, visible_threshold=1, conf_threshold=0.15)
    P.SHARD_VIEW_UPLOAD = True
    pm_f64 = P.PMVO(cameras_from_scene(sc), sc.ref_depths(), Ori, Conf, sc.ref_masks(), device=dev,
                    image_size=[sc.H, sc.W], patch_size=7, visible_threshold=1, conf_threshold=0.15)
    maps_ok = (torch.equal(pm.mapC, pm_full.mapC) and torch.equal(pm.mapP, pm_full.mapP)
               and torch.equal(pm_f64.mapC, pm_full.mapC) and torch.equal(pm_f64.mapP, pm_full.mapP))
    if dist.get_rank() == 0:
        print(f"sharded maps identical: {maps_ok}")
    del pm_full, pm_f64
    c = torch.from_numpy(cand).to(dev).float()
    multi = pipeline.pmvo_job_device(pm, c, 0.025)
    pipeline._FORCE_SINGLE = True
    single = pipeline.pmvo_job_device(pm, c, 0.025)
    pipeline._FORCE_SINGLE = False
    ok = maps_ok
    for k in ("surface", "filter", "select_o", "min_loss", "high_conf", "refine_o", "refine_loss", "fu_ori", "volume"):
        same = torch.equal(multi[k], single[k])
        ok &= same
        if dist.get_rank() == 0:
            print(f"{k:12s} identical: {same}")
    # the chunk-ordered sweep with many small chunks (the job above has one or two): in "peer" mode every rank owns every
    # world-th block of 64 points and the finished points travel through symmetric memory (mh_refine_sweep_dist)
    pts, o0, l0 = single["select_p"], single["select_o"], single["min_loss"]
    for sub in (257, 1000):
        o_m, l_m = pipeline.refine_stage(pm, pts, o0, l0, sub_num=sub)
        pipeline._FORCE_SINGLE = True
        o_s, l_s = pipeline.refine_stage(pm, pts, o0, l0, sub_num=sub)
        pipeline._FORCE_SINGLE = False
        same = torch.equal(o_m, o_s) and torch.equal(l_m, l_s)
        ok &= same
        if dist.get_rank() == 0:
            print(f"refine sweep ({pipeline._sweep_mode(dist, dev)}, {pts.size(0)} points, chunks of {sub}) identical: {same}; "
                  f"{int((o_s != o0).any(1).sum().item())} orientations updated")
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if dist.get_rank() == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", "world", dist.get_world_size(),
              "fusion", os.environ.get("MH_FUSE_DIST", "replicated"), "sweep", os.environ.get("MH_SWEEP_DIST", "peer"),
              "occupied voxels", int(multi["volume"][..., 3].sum().item()))
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
