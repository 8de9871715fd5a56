"""ctypes binding of libmonohair_b200.so (include/monohair_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MH_LIB") or os.path.join(HERE, "libmonohair_b200.so")   # MH_LIB: tuning experiments only

MH_CAM_STRIDE = 32
MH_TOPK = 20
MH_NUM_BASE = 10


class MhViews(C.Structure):
    _fields_ = [("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("P", C.c_int32),
                ("mapC", C.c_void_p), ("mapP", C.c_void_p), ("cam", C.c_void_p)]


class MonoHairError(RuntimeError):
    pass


_lib = None
p, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
VP = C.POINTER(MhViews)

_SIGS = {
    "mh_last_error": (C.c_char_p, []),
    "mh_version": (C.c_int, []),
    "mh_launch_count": (i64, []),
    "mh_views_pack_camera_host": (C.c_int, [p, p, p, p]),
    "mh_views_pack": (C.c_int, [p, i32, i32, i32, i32, p, i32, p, p, p, i32, p, p]),
    "mh_views_pack_f64": (C.c_int, [p, i32, i32, i32, i32, p, i32, p, p, p, i32, p, p]),
    "mh_views_pack_u8": (C.c_int, [p, i32, i32, i32, i32, p, i32, p, p, p, p, p, p, p, p]),
    "mh_filter_count": (C.c_int, [p, VP, p, i64, f32, f32, p]),
    "mh_filter_decide": (C.c_int, [p, p, i64, p, p]),
    "mh_visible_count": (C.c_int, [p, VP, p, i64, f32, p]),
    "mh_head_count": (C.c_int, [p, VP, p, i64, f32, p]),
    "mh_head_decide": (C.c_int, [p, p, p, p, i64, f64, f64, p]),
    "mh_centre_gather": (C.c_int, [p, VP, p, i64, p, p, p, p, p]),
    "mh_refine_chunks": (C.c_int, [p, VP, p, p, i32, p, i64, i64, f32, p, p, p, i64]),
    "mh_refine_chunks_workspace_bytes": (i64, [i64, i64]),
    "mh_refine_sweep_workspace_bytes": (i64, [i64, i64]),
    "mh_refine_sweep": (C.c_int, [p, p, p, i32, i64, i64, p, p, p, i64]),
    "mh_refine_finish": (C.c_int, [p, p, p, i64, p]),
    "mh_refine_sweep_dist_block": (i64, []),
    "mh_refine_sweep_dist_local_count": (i64, [i64, i32, i32]),
    "mh_refine_sweep_dist": (C.c_int, [p, p, p, i32, i64, i64, i32, i32, p, p, f64, i32, p, i64, p]),
    "mh_refine_update": (C.c_int, [p, p, p, p, i64, p, p]),
    "mh_pmvo_optimize_workspace_bytes": (i64, [VP, i64]),
    "mh_pmvo_optimize": (C.c_int, [p, VP, p, i64, p, i32, f32, p, p, p, p, p, p, p, p, p, i64]),
    "mh_pmvo_refine_loss": (C.c_int, [p, VP, p, p, i64, f32, p]),
    "mh_knn_workspace_bytes": (i64, [i64, i64, i32]),
    "mh_knn": (C.c_int, [p, p, i64, p, i64, i32, p, f64, p, p, i64]),
    "mh_knn_q64": (C.c_int, [p, p, i64, p, i64, i32, p, f64, p, p, i64]),
    "mh_nn_dist": (C.c_int, [p, p, i64, p, i64, p]),
    "mh_medoid_gather": (C.c_int, [p, p, p, i64, i32, p, p]),
    "mh_voxel_fuse_workspace_bytes": (i64, [i64, i32, i32, i32]),
    "mh_voxel_fuse": (C.c_int, [p, p, p, p, i64, p, f64, i32, i32, i32, p, p, p, p, i64]),
    "mh_voxel_fuse_winners": (C.c_int, [p, p, p, p, i64, p, f64, i32, i32, i32, p, i64, p, p, p, p, i64]),
    "mh_voxel_scatter": (C.c_int, [p, p, i64, i32, i32, i32, p, i32]),
    "mh_voxel_fuse_plane_bytes": (i64, [i32, i32, i32]),
    "mh_voxel_fuse_plane_init": (C.c_int, [p, p, i32, i32, i32]),
    "mh_voxel_fuse_max_points": (C.c_int, [p, p]),
    "mh_voxel_overwrite": (C.c_int, [p, p, p, i64, p, f64, i32, i32, i32, p, p]),
    "mh_volume_to_mat": (C.c_int, [p, p, i32, i32, i32, p, p]),
    "mh_volume_from_mat": (C.c_int, [p, p, p, i32, i32, i32, p]),
    "mh_trace_count": (C.c_int, [p, p, i32, i32, i32, p, i64, f32, i32, p, p]),
    "mh_trace_write": (C.c_int, [p, p, i32, i32, i32, p, i64, f32, i32, p, p, p, i32, p]),
    "mh_trace_from_scalp": (C.c_int, [p, p, i32, i32, i32, p, p, i64, f32, i32, i32, p, p]),
    "mh_accept_strands": (C.c_int, [p, p, p, p, p, i64, i32, i32, i32, i32, p, p]),
    "mh_accept_strands_workspace_bytes": (i64, [i64, i64]),
    "mh_accept_strands_ws": (C.c_int, [p, p, p, p, p, i64, i64, i32, i32, i32, p, p, p, i64]),
    "mh_sample_mark_cells": (C.c_int, [p, p, i64, p, f64, i32, i32, i32, p]),
    "mh_sample_cells": (C.c_int, [p, p, i64, i32, p, p, f64, p]),
    "mh_render_depth": (C.c_int, [p, p, i64, p, i64, p, i32, i32, p, p, i32]),
    "mh_connect_find": (C.c_int, [p, p, p, p, i64, f64, f64, p, p]),
    "mh_strand_occupancy": (C.c_int, [p, p, p, p, i64, p, p, i32, i32, i32, i32, p]),
    "mh_smooth_strands_workspace_bytes": (i64, [i64]),
    "mh_smooth_strands": (C.c_int, [p, p, p, p, i64, f64, f64, p, p, i64, i64]),
    "mh_gabor_workspace_bytes": (i64, [i32, i32, i32]),
    "mh_gabor_orientation": (C.c_int, [p, p, i32, i32, p, i32, i32, f32, f32, p, p, p, p, i64]),
    "mh_gabor_tc_bank_bytes": (i64, []),
    "mh_gabor_orientation_tc": (C.c_int, [p, p, i32, i32, p, i32, f32, f32, p, p, p, p, i64]),
    "mh_filterbank_wrap_f64": (C.c_int, [p, p, i32, i32, p, i32, i32, p]),
    "mh_dog_f64": (C.c_int, [p, p, i32, i32, p, i32, p, i32, p, p]),
    "mh_debug_topk_host": (C.c_int, [p, i32, i32, p, p]),
    "mh_debug_div2_check": (C.c_int, [p, p, p, p, i64, p]),
}


def exported_symbols():
    return list(_SIGS.keys())


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MonoHairError(f"{LIB_PATH} is missing: run `python -m monohair_b200.build` "
                                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise MonoHairError(f"{what} failed ({rc}): {lib().mh_last_error().decode()}")


def ptr(t):
    """device/host pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dev_tensor(t, dtype, device):
    assert t.is_cuda and t.device == torch.device(device) and t.dtype == dtype and t.is_contiguous(), \
        f"expected contiguous {dtype} tensor on {device}, got {t.dtype} on {t.device}"
    return t
