"""Drop-in entry point: `python HairGrow.py --yaml=configs/reconstruct/<case>` -- the generate_segments phase of the
reference's HairGrow.py (:876-919) on monohair_b200's trace kernels; writes scalp_segment.hair,
scalp_segment_smooth.hair and num_root.npy.
The connect / smooth phases (HairGrow.py:925-976) are "next" rows of SURVEY.md §8f and are not part of this path."""
import os
import sys

import numpy as np
import torch

from monohair_b200 import options
from monohair_b200.hairgrow import (HairGrowing, load_strand, points_to_voxel, save_hair_strands, smooth_strands,  # noqa: F401
                                    voxel_to_points)
from monohair_b200.pmvo_utils import read_obj, read_obj_normals, sample_points_uniformly  # noqa: F401


def config_parser():
    """HairGrow.py:837-873."""
    opt_cmd = options.parse_arguments(sys.argv[1:])
    args = options.set(opt_cmd=opt_cmd)
    args.output_path = os.path.join(args.data.root, args.data.case, args.output_root, args.name)
    os.makedirs(args.output_path, exist_ok=True)
    options.save_options_file(args)
    args.data.root = os.path.join(args.data.root, args.data.case)
    args.bbox_min = np.array(args.bbox_min)
    args.bust_to_origin = np.array(args.bust_to_origin)
    for key in ("strands_path", "bust_path", "scalp_path"):
        args.data[key] = os.path.join(args.data.root, args.data[key])
    suffix = '_diffusion' if args.scalp_diffusion else ''
    args.image_camera_path = os.path.join(args.data.root, args.image_camera_path)
    args.save_path = os.path.join(args.output_path, 'full' if args.PMVO.infer_inner else 'refine')
    args.data.Occ3D_path = os.path.join(args.save_path, 'Occ3D{}.mat'.format(suffix))
    args.data.Ori3D_path = os.path.join(args.save_path, 'Ori3D{}.mat'.format(suffix))
    return args


def main():
    args = config_parser()
    v, f, vn = read_obj_normals(args.data.scalp_path)            # open3d read_triangle_mesh: `vn` records as vertex normals
    scalp_points, scalp_normals = sample_points_uniformly(v, f, 60000, with_normals=True, vertex_normals=vn)   # use_triangle_normal=False
    scalp_points += args.bust_to_origin
    scalp_points = torch.from_numpy(scalp_points).to(args.device)
    scalp_normals = torch.from_numpy(scalp_normals).to(args.device)
    scalp_normals = scalp_normals / torch.linalg.norm(scalp_normals, 2, -1, keepdims=True)
    scalp_points = points_to_voxel(scalp_points)
    scalp_normals[:, 1:] *= -1
    scalp_normals = scalp_normals.type(torch.float32)
    scalp_points = scalp_points.type(torch.float32)
    solver = HairGrowing(args.data.Occ3D_path, args.data.Ori3D_path, device=args.device, image_size=args.data.image_size)
    if args.HairGenerate.generate_segments:
        strands, num_root = solver.GenerateGuideStrandFromScalp(scalp_points, scalp_normals, None, args.HairGenerate.grow_threshold)
        strands = solver.VoxelToWorld(strands, args.bust_to_origin)
        save_hair_strands(os.path.join(args.save_path, 'scalp_segment.hair'), strands)
        strands = smooth_strands(strands, 4.0, 2.0, device=args.device)                       # HairGrow.py:914-916
        save_hair_strands(os.path.join(args.save_path, 'scalp_segment_smooth.hair'), strands)
        np.save(args.save_path + '/num_root.npy', np.array(num_root))
    else:
        num_root = int(np.load(args.save_path + '/num_root.npy'))
    if args.HairGenerate.connect_segments:                                                      # HairGrow.py:925-950
        segment, points = load_strand(os.path.join(args.save_path, 'scalp_segment.hair'))
        strands, beg = [], 0
        for i, seg in enumerate(segment):
            strand = points[beg:beg + seg]
            if i >= num_root:
                strand += args.bust_to_origin
            strands.append(strand)
            beg += seg
        solver.strands = strands
        connected = solver.find_connect_info(strands[num_root:], args.HairGenerate.connect_threshold,
                                             args.HairGenerate.connect_dot_threshold, solver.occ)
        new_strands = strands[:num_root] + [ss - args.bust_to_origin for ss in connected]
        new_strands = smooth_strands(new_strands, 4.0, 2.0, device=args.device)
        save_hair_strands(os.path.join(args.save_path, 'strands.hair'), new_strands)
    if args.HairGenerate.connect_scalp:                                                         # HairGrow.py:952-976
        segment, points = load_strand(os.path.join(args.save_path, 'strands.hair'))
        strands, beg = [], 0
        for seg in segment:
            strands.append(points[beg:beg + seg])
            beg += seg
        strands = solver.WorldToVoxel(strands, args.bust_to_origin)
        connect_strands = solver.connect_to_scalp(strands, num_root, args.HairGenerate.out_ratio, bool(args.PMVO.infer_inner))
        out = []
        for ss in connect_strands:
            ss = voxel_to_points(torch.from_numpy(ss.copy())).cpu().numpy()
            ss -= args.bust_to_origin
            out.append(ss)
        out = smooth_strands(out, 4.0, 2.0, device=args.device)
        save_hair_strands(os.path.join(args.save_path, 'connected_strands.hair'), out)


if __name__ == '__main__':
    main()
