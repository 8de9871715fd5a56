"""Drop-in entry point: `python PMVO.py --yaml=configs/reconstruct/<case> [--PMVO.infer_inner --PMVO.optimize=]`
(same CLI, YAML keys and output files as the reference's PMVO.py:805-880), running on monohair_b200's CUDA kernels.
Under torchrun (one process per GPU) the point-parallel stages are sharded and rank 0 writes the files."""
import os

import numpy as np
import torch

from monohair_b200 import pmvo as _impl
from monohair_b200.camera import load_cam, parsing_camera
from monohair_b200.pmvo import PMVO, config_parser, filter_negative_points, optimize, refine  # noqa: F401 (reference names)
from monohair_b200.pmvo_utils import (Load_Ori_And_Conf, load_bust, load_colmap_points, load_depth, load_mask, read_obj)


def main():
    print('Run PMVO...')
    args = config_parser()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # one process per GPU: every rank loads the capture, rank 0 samples the candidates (the sampler draws random
        # numbers) and broadcasts them, the point-parallel stages are sharded (monohair_b200/pipeline.py) and rank 0
        # alone writes the files; the other ranks read them back behind a barrier where the reference re-reads its own
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        args.device = f"cuda:{local}"
    rank = int(os.environ.get("RANK", "0"))
    _impl.device = args.device
    _impl.args = args
    vertices, faces, _ = load_bust(args.data.bust_path)
    vertices += args.bust_to_origin
    _impl.bust_tree = vertices
    scalp_vertices, _ = read_obj(os.path.join(args.data.root, 'ours/scalp_tsfm.obj'))
    scalp_vertices += args.bust_to_origin
    _impl.scalp_tree = scalp_vertices
    _impl.scalp_max = np.max(scalp_vertices, axis=0)

    camera = parsing_camera(load_cam(args.image_camera_path), os.path.join(args.data.root, 'capture_images'))
    print('num of view:', len(camera))
    depths = load_depth(camera, args.data.depth_path)
    Ori, Conf = Load_Ori_And_Conf(camera, args.data.Ori2D_path, args.data.Conf_path)
    masks = load_mask(camera, args.data.mask_path)
    pmvo = PMVO(camera, depths, Ori, Conf, masks, device=args.device, image_size=args.data.image_size,
                patch_size=args.PMVO.patch_size, visible_threshold=args.PMVO.visible_threshold,
                conf_threshold=args.PMVO.conf_threshold)
    del depths, Ori, Conf, masks

    if args.PMVO.optimize:
        print('load raw mesh...')
        points = None
        if rank == 0:
            points = load_colmap_points(args.data.raw_points_path, args.bbox_min, args.bust_to_origin, 0.005 / 4,
                                        [512, 512, 384], True, args.PMVO.num_sample_per_grid, device=args.device)
        points = _impl.broadcast_array(points, args.device)
        raw_points = points.copy()
        print('total points:', points.shape[0])
        if args.PMVO.filter_point:
            print('filter low conf points...')
            surface_index, surface_points, filter_index = filter_negative_points(points, pmvo, args)
            n_cov = surface_index.shape[0]
            points = surface_points
            if rank == 0:
                os.makedirs(args.save_root, exist_ok=True)
                np.save(os.path.join(args.save_root, 'surface.npy'), raw_points[:n_cov][surface_index])
                np.save(os.path.join(args.save_root, 'filter_unvisible.npy'), raw_points[:n_cov][filter_index])
        _impl.Num_points = points.shape[0]
        print('process points:', _impl.Num_points)
        optimize(points, pmvo, args)                 # rank 0 writes optimize/*.npy, then a barrier
        select_points = np.load(args.save_root + '/select_p.npy')
        select_ori = np.load(args.save_root + '/select_o.npy')
        min_loss = np.load(args.save_root + '/min_loss.npy')
        filter_unvisible_points = np.load(args.save_root + '/filter_unvisible.npy')
        refine(select_points, select_ori, min_loss, pmvo, filter_unvisible_points, args, infer_inner=False,
               threshold=args.PMVO.threshold, genrate_ori_only=False)
    else:
        select_points = np.load(args.save_root + '/select_p.npy')
        select_ori = np.load(args.save_root + '/select_o.npy')
        min_loss = np.load(args.save_root + '/min_loss.npy')
        filter_unvisible_points = np.load(args.save_root + '/filter_unvisible.npy')
        refine(select_points, select_ori, min_loss, pmvo, filter_unvisible_points, args, infer_inner=args.PMVO.infer_inner,
               threshold=args.PMVO.threshold, genrate_ori_only=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
