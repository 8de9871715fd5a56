"""Drop-in module name of the reference (Utils/PMVO_utils.py): re-exports the hot-path functions."""
from monohair_b200.pmvo_utils import *  # noqa: F401,F403
from monohair_b200.pmvo_utils import (Load_Ori_And_Conf, SamplePointsAroundmesh, compute_points_similarity,  # noqa: F401
                                      get_ground_truth_3D_occ, get_ground_truth_3D_ori, load_colmap_points, load_depth,
                                      load_mask, p2v, points_to_voxel, voxel_to_points)
