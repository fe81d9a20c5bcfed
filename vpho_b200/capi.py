"""ctypes binding of the C ABI declared in include/vpho_b200.h.

`lib()` is the only loader the product uses: it loads `vpho_b200/csrc/libvpho_b200.so` (sm_100a) and raises when the
library is missing or no CUDA device is visible -- there is no CPU path and no fallback.  `Library(path)` exists so
tests can bind an explicitly named binary (e.g. the SIMT emulation build under tests/emu/_build/).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvpho_b200.so")

c_void_p, c_int, c_float, c_double, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t


class SampleArgs(C.Structure):
    """`vpho_sample_args` (include/vpho_b200.h): the per-sampler arguments of vpho_sample_begin."""
    _fields_ = [
        ("denoiser", c_void_p), ("feat", c_void_p), ("n_rows", c_int), ("rows_per_feat", c_int), ("init_x", c_void_p),
        ("T0", c_double), ("eps", c_double), ("t_eval", c_void_p), ("n_eval", c_int),
        ("rtol", c_double), ("atol", c_double), ("max_step", c_double), ("num_steps", c_int),
        ("xs", c_void_p), ("x", c_void_p), ("counters", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("xs_f32", c_void_p),
    ]


class HoiArgs(C.Structure):
    _fields_ = [
        ("bs", c_int), ("S", c_int), ("topk_hand", c_int), ("topk_obj", c_int), ("phy_topk", c_int),
        ("cam_intrinsic", c_void_p), ("root_joint_flip", c_void_p), ("root_joint", c_void_p),
        ("is_right", c_void_p), ("is_grasped", c_void_p), ("force_local", c_void_p),
        ("hand_pose_diff", c_void_p), ("hand_pose_reg", c_void_p), ("hand_shape", c_void_p),
        ("hand_heatmap", c_void_p), ("hand_bbox", c_void_p), ("obj_pose6d", c_void_p),
        ("obj_heatmap", c_void_p), ("obj_bbox", c_void_p), ("obj_id", c_void_p),
        ("obj_agg_6d", c_void_p), ("pose6d_candidate", c_void_p), ("agg_obj_vert", c_void_p),
        ("hand_agg_mano", c_void_p), ("hand_agg_vert", c_void_p), ("hand_agg_joint", c_void_p),
        ("dbg_hand_score", c_void_p), ("dbg_hand_topk", c_void_p), ("dbg_cascade_pose", c_void_p),
        ("dbg_obj_score", c_void_p), ("dbg_obj_topk", c_void_p), ("dbg_finger_score", c_void_p),
        ("dbg_finger_topk", c_void_p), ("dbg_force_point", c_void_p), ("dbg_force_global", c_void_p),
    ]


class NamedTensor(C.Structure):
    """`vpho_named_tensor` (include/vpho_b200.h): one entry of a reference state dict."""
    _fields_ = [("name", C.c_char_p), ("data", c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class HeadsArgs(C.Structure):
    """`vpho_heads_args` (include/vpho_b200.h)."""
    _fields_ = [("bs", C.c_int32), ("roi_size", C.c_int32)] + [(k, c_void_p) for k in (
        "hf_hr", "of_or_rect", "hf_hr_rect", "bbox_hand", "bbox_hand_rect", "bbox_obj", "bbox_obj_rect", "is_right", "gravity",
        "hand_heatmap", "obj_heatmap", "encoding_hand", "encoding_obj", "mano_pose", "mano_shape", "force_local", "force_scale",
        "force_weight", "CoM", "enc_phy_hand", "enc_phy_obj")] + [("flags", C.c_int32)]


class EvalRecordArgs(C.Structure):
    """`vpho_eval_record_args` (include/vpho_b200.h)."""
    _fields_ = [("bs", c_int), ("S", c_int)] + [(k, c_void_p) for k in (
        "agg_hand_joint", "agg_hand_vert", "cand_hand_joint", "cand_hand_vert", "reg_hand_joint", "reg_hand_vert",
        "agg_obj_6d", "cand_obj_6d", "root_joint", "is_right", "gt_joint", "gt_vert", "gt_obj_rt", "cam_intr", "obj_id",
        "out")]


class JointScoresArgs(C.Structure):
    """`vpho_joint_scores_args` (include/vpho_b200.h)."""
    _fields_ = [("bs", c_int), ("n", c_int), ("n_joints", c_int)] + [(k, c_void_p) for k in (
        "joint", "root_joint", "cam_intrinsic", "bbox", "heatmap", "heat", "dist2d")]


class HandLevelArgs(C.Structure):
    """`vpho_hand_level_args` (include/vpho_b200.h)."""
    _fields_ = [("bs", c_int), ("n", c_int), ("K", c_int), ("n_joints", c_int), ("score", c_void_p), ("pose", c_void_p),
                ("joint", c_void_p), ("observe_index", c_void_p), ("n_observe", c_int), ("fuse_index", c_void_p),
                ("n_fuse", c_int), ("independent", c_int), ("is_weight", c_int), ("fuse_kind", c_int), ("write_back", c_int),
                ("val", c_void_p), ("topk", c_void_p), ("fused", c_void_p)]


class ObjSelectArgs(C.Structure):
    """`vpho_obj_select_args` (include/vpho_b200.h)."""
    _fields_ = [("bs", c_int), ("n", c_int), ("K", c_int)] + [(k, c_void_p) for k in (
        "pose6d", "root_joint", "cam_intrinsic", "bbox", "heatmap", "is_right", "obj_id")] + [("is_weight", c_int), ("score_kind", c_int)] + [
        (k, c_void_p) for k in ("topk_in", "topk", "weight", "fused")]


_SIGNATURES = {
    "vpho_version": (c_int, []),
    "vpho_launch_count": (C.c_ulonglong, []),
    "vpho_measure_peaks": (c_int, [C.POINTER(c_float), C.POINTER(c_float), c_int, c_void_p]),
    "vpho_set_pdl": (c_int, [c_int]),
    "vpho_profile_reserve": (c_int, [c_int]),
    "vpho_profile_enable": (c_int, [c_int]),
    "vpho_profile_collect": (c_int, [c_int, C.POINTER(c_double), C.POINTER(c_int)]),
    "vpho_profile_collect_list": (c_int, [c_int, C.POINTER(c_double), C.POINTER(c_int), c_void_p, c_int]),
    "vpho_mano_create": (c_int, [c_void_p] * 5 + [C.POINTER(c_void_p)]),
    "vpho_mano_destroy": (c_int, [c_void_p]),
    "vpho_mano_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vpho_mano_forward_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "vpho_denoiser_create": (c_int, [c_int] + [c_void_p] * 11 + [C.POINTER(c_void_p)]),
    "vpho_denoiser_create_ex": (c_int, [c_int] + [c_void_p] * 11 + [c_int, C.POINTER(c_void_p)]),
    "vpho_denoiser_destroy": (c_int, [c_void_p]),
    "vpho_score_eval": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "vpho_sample_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vpho_sample_begin": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_double, c_double, c_void_p, c_int,
                                  c_double, c_double, c_double, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "vpho_sample_continue": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vpho_sample_finish": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vpho_sample_pair_begin": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "vpho_sample_pair_continue": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "vpho_sample_pair_finish": (c_int, [c_void_p, c_void_p, c_void_p]),
    "vpho_rot6d_to_axis_angle": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "vpho_postprocess_hand": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vpho_postprocess_hand_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vpho_assets_create": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   C.POINTER(c_void_p)]),
    "vpho_assets_destroy": (c_int, [c_void_p]),
    "vpho_object_points": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p]),
    "vpho_force_anchors": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vpho_anchor_contact": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vpho_vertex_contact": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vpho_force_eval": (c_int, [c_void_p] * 9 + [c_int, c_int] + [c_void_p] * 5),
    "vpho_force_optimize_workspace_bytes": (c_size_t, [c_int]),
    "vpho_force_optimize": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_float] + [c_void_p] * 6 + [c_size_t, c_void_p]),
    "vpho_pose_metrics": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p]),
    "vpho_hand_metrics": (c_int, [c_void_p] * 4 + [c_int, c_void_p, c_void_p]),
    "vpho_objmetrics_create": (c_int, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                       C.POINTER(c_void_p)]),
    "vpho_objmetrics_destroy": (c_int, [c_void_p]),
    "vpho_object_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "vpho_eval_record_workspace_bytes": (c_size_t, [c_int]),
    "vpho_eval_record": (c_int, [c_void_p, c_void_p, C.POINTER(EvalRecordArgs), c_void_p, c_size_t, c_void_p]),
    "vpho_hand_pa_metrics": (c_int, [c_void_p] * 4 + [c_int, c_void_p, c_void_p]),
    "vpho_heads_create": (c_int, [C.POINTER(NamedTensor), c_int, C.POINTER(c_void_p)]),
    "vpho_heads_destroy": (c_int, [c_void_p]),
    "vpho_heads_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "vpho_heads_forward": (c_int, [c_void_p, C.POINTER(HeadsArgs), c_void_p, c_size_t, c_void_p]),
    "vpho_heads_dims": (c_int, [c_void_p, C.POINTER(C.c_int32)]),
    "vpho_heads_overflow": (c_int, [c_void_p, C.POINTER(C.c_int32), c_void_p]),
    "vpho_hoi_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vpho_hoi_aggregate": (c_int, [c_void_p, c_void_p, C.POINTER(HoiArgs), c_void_p, c_size_t, c_void_p]),
    "vpho_joint_scores": (c_int, [C.POINTER(JointScoresArgs), c_void_p, c_size_t, c_void_p]),
    "vpho_hand_level": (c_int, [C.POINTER(HandLevelArgs), c_void_p]),
    "vpho_quat_average_all": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vpho_obj_select": (c_int, [c_void_p, C.POINTER(ObjSelectArgs), c_void_p, c_size_t, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class VphoError(RuntimeError):
    pass


class Library:
    """A bound shared library exporting the vpho_b200 C ABI."""

    def __init__(self, path: str, strict: bool = True):
        if not os.path.exists(path):
            raise VphoError(f"vpho_b200 native library not found: {path} (build it with `python -m vpho_b200.build`)")
        self.path = path
        self.c = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            if not strict and not hasattr(self.c, name):
                continue
            fn = getattr(self.c, name)
            fn.restype = res
            fn.argtypes = args

    def check(self, status: int, what: str) -> None:
        if status != 0:
            reason = {-1: "invalid argument", -2: "kernel launch failure", -3: "allocation failure"}.get(status, "?")
            raise VphoError(f"{what} failed with status {status} ({reason})")


_default: Optional[Library] = None


def lib() -> Library:
    """The product library.  Fails loudly instead of falling back (north_star: no CPU fallback)."""
    global _default
    if _default is None:
        if not torch.cuda.is_available():
            raise VphoError("vpho_b200 needs a CUDA device (built for sm_100a); there is no CPU implementation")
        _default = Library(LIB_PATH)
    return _default


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> c_void_p:
    if t is None:
        return c_void_p(None)
    if dtype is not None and t.dtype != dtype:
        raise VphoError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise VphoError("tensor must be contiguous")
    return c_void_p(t.data_ptr())


def stream_of(t: torch.Tensor) -> c_void_p:
    if t.is_cuda:
        return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return c_void_p(None)


def host_ptr(a) -> c_void_p:
    """Pointer to a C-contiguous numpy array that the caller keeps alive."""
    return c_void_p(a.ctypes.data)
