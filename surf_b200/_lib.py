"""ctypes binding of include/surf_b200.h (the C-ABI of csrc/libsurf_b200.so).

No torch types cross the boundary: tensors are passed as ``data_ptr()`` integers plus sizes, the
stream as ``torch.cuda.current_stream().cuda_stream``.  There is no CPU fallback: if the library
is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

MAX_LEVELS = 4
MAX_VIEWS = 8
MAX_STAGES = 4
SDF_LAYERS = 7
ABI_VERSION = 4
COLOR_SERIAL, COLOR_OVERLAP, COLOR_FUSED = 0, 1, 2   # surf_render_cfg.color_path

# MLP kernel family, chosen per call (include/surf_b200.h SURF_MLP_*)
MLP_FFMA = 0        # fp32 CUDA-core kernels: the parity anchor
MLP_TC = 1          # tcgen05 / TMEM kernels, fp16 hi/lo 3-MMA split: fp32-grade, the default of the Python mirror
MLP_TC_FAST = 4     # tcgen05, one fp16 MMA per product: opt-in reduced precision (1e-2)
MLP_MODES = (MLP_FFMA, MLP_TC, MLP_TC_FAST)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsurf_b200.so")

EXPORTS = [
    "surf_version", "surf_last_error", "surf_launch_count", "surf_timing_enable", "surf_timing_read",
    "surf_scene_create", "surf_scene_destroy", "surf_scene_get_stats", "surf_scene_update_volume",
    "surf_scene_set_views",
    "surf_net_create", "surf_net_destroy",
    "surf_render_workspace_bytes", "surf_sample_rays", "surf_render_core", "surf_render_rays",
    "surf_sdf_points", "surf_sdf_grid", "surf_sdf_full", "surf_sdf_smooth",
    "surf_mask_dilate", "surf_mesh_vertex_visibility", "surf_mesh_raster_workspace_bytes", "surf_mesh_first_hits",
    "surf_mesh_components_workspace_bytes", "surf_mesh_components",
    "surf_mt19937_skip",
    "surf_fpn_conv_blocks", "surf_fpn_deconv_blocks", "surf_fpn_conv3x3", "surf_fpn_deconv3x3s2", "surf_fpn_finish_stats",
    "surf_fpn_norm_relu_add",
    "surf_point_mask", "surf_lookup_sparse", "surf_lookup_feature", "surf_blend", "surf_point_flags",
    "surf_tc_selftest",
    "surf_mc_workspace_bytes", "surf_mc_count", "surf_mc_emit",
    "surf_extras_workspace_bytes", "surf_render_extras", "surf_depth_map",
    "surf_volume_back_proj", "surf_volume_depth_filter", "surf_volume_upsample2x", "surf_scene_create_sparse",
]


class SceneInputs(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32),
        ("feat_ch", C.c_int32),
        ("dim", C.c_int32 * MAX_LEVELS),
        ("n_vox", C.c_int64 * MAX_LEVELS),
        ("d_volumes", C.c_void_p * MAX_LEVELS),
        ("d_sparse_idx", C.c_void_p * MAX_LEVELS),
        ("d_mask_volumes", C.c_void_p * MAX_LEVELS),
        ("d_matching_volume", C.c_void_p),
        ("match_dim", C.c_int32),
        ("n_views", C.c_int32),
        ("img_h", C.c_int32),
        ("img_w", C.c_int32),
        ("n_feat_levels", C.c_int32),
        ("d_imgs", C.c_void_p),
        ("d_features", C.c_void_p * 4),
        ("h_intrs", C.c_void_p),
        ("h_w2cs", C.c_void_p),
        ("h_c2ws", C.c_void_p),
    ]


class SceneViews(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32),
        ("img_h", C.c_int32),
        ("img_w", C.c_int32),
        ("n_feat_levels", C.c_int32),
        ("d_imgs", C.c_void_p),
        ("d_features", C.c_void_p * 4),
        ("h_intrs", C.c_void_p),
        ("h_w2cs", C.c_void_p),
        ("h_c2ws", C.c_void_p),
    ]


class SceneStats(C.Structure):
    _fields_ = [
        ("bytes_index", C.c_int64), ("bytes_volumes", C.c_int64), ("bytes_masks", C.c_int64),
        ("bytes_matching", C.c_int64), ("bytes_images", C.c_int64),
        ("n_vox", C.c_int64 * MAX_LEVELS),
    ]


class NetInputs(C.Structure):
    _fields_ = [
        ("n_lin", C.c_int32),
        ("in_dim", C.c_int32 * SDF_LAYERS),
        ("out_dim", C.c_int32 * SDF_LAYERS),
        ("h_weight_v", C.c_void_p * SDF_LAYERS),
        ("h_weight_g", C.c_void_p * SDF_LAYERS),
        ("h_bias", C.c_void_p * SDF_LAYERS),
        ("multires", C.c_int32),
        ("skip_layer", C.c_int32),
        ("feat_channels", C.c_int32),
        ("scale", C.c_float),
        ("h_blend_w", C.c_void_p * 11),
        ("h_blend_b", C.c_void_p * 11),
        ("blend_s", C.c_float),
        ("d_feature", C.c_int32),
        ("variance", C.c_float),
    ]


class RenderCfg(C.Structure):
    _fields_ = [
        ("n_stages", C.c_int32),
        ("n_samples", C.c_int32 * MAX_STAGES),
        ("sample_ranges", C.c_float * MAX_STAGES),
        ("n_depth", C.c_int32),
        ("perturb", C.c_int32),
        ("cos_anneal_ratio", C.c_float),
        ("chunk_rays", C.c_int32),
        ("d_lin_tables", C.c_void_p),
        ("mlp_mode", C.c_int32),
        ("color_path", C.c_int32),
    ]


RENDER_OUTPUT_FIELDS = [
    "d_color_fine", "d_render_depth", "d_sdf_depth", "d_normal", "d_val_normal", "d_weights", "d_weight_sum",
    "d_weight_max", "d_valid_mask", "d_inside_sphere", "d_mid_inside_sphere", "d_mid_z_vals", "d_gradients",
    "d_sdf", "d_gradient_error_sums", "d_point_flags", "d_point_color", "d_point_views", "d_prev_idx", "d_alpha",
    "d_z_cross", "d_z_max",
]


class ExtrasParams(C.Structure):
    """surf_extras_params (include/surf_b200.h)."""
    _fields_ = [("R0t", C.c_float * 9), ("t0", C.c_float * 3), ("K0", C.c_float * 9), ("K0inv", C.c_float * 9),
                ("Ksrc", (C.c_float * 9) * MAX_VIEWS), ("Rrel", (C.c_float * 9) * MAX_VIEWS),
                ("RC", (C.c_float * 3) * MAX_VIEWS), ("n_src", C.c_int32), ("patch_size", C.c_int32)]


class VolumeViews(C.Structure):
    """surf_volume_views (include/surf_b200.h)."""
    _fields_ = [("n_views", C.c_int32), ("norm_h", C.c_int32), ("norm_w", C.c_int32), ("h_w2cs", C.c_void_p),
                ("h_intrs", C.c_void_p)]


class SceneSparseInputs(C.Structure):
    """surf_scene_sparse_inputs (include/surf_b200.h)."""
    _fields_ = [("n_levels", C.c_int32), ("feat_ch", C.c_int32), ("dim", C.c_int32 * MAX_LEVELS),
                ("n_vox", C.c_int64 * MAX_LEVELS), ("d_coords", C.c_void_p * MAX_LEVELS),
                ("d_volumes", C.c_void_p * MAX_LEVELS), ("d_logits", C.c_void_p * MAX_LEVELS)]


class DepthMapParams(C.Structure):
    """surf_depth_map_params (include/surf_b200.h)."""
    _fields_ = [("Kinv", C.c_float * 9), ("R", C.c_float * 9), ("C", C.c_float * 3), ("Rinv2", C.c_float * 3),
                ("near", C.c_float), ("far", C.c_float), ("ratio", C.c_float * 2), ("n_windows", C.c_int32),
                ("n_samples", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32)]


class RenderOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in RENDER_OUTPUT_FIELDS]


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    P = C.POINTER
    lib.surf_version.restype = C.c_int
    lib.surf_version.argtypes = []
    lib.surf_last_error.restype = C.c_char_p
    lib.surf_last_error.argtypes = []
    lib.surf_launch_count.restype = i64
    lib.surf_launch_count.argtypes = []
    lib.surf_timing_enable.restype = C.c_int
    lib.surf_timing_enable.argtypes = [i32]
    lib.surf_timing_read.restype = C.c_int
    lib.surf_timing_read.argtypes = [P(C.c_double), P(i64)]
    lib.surf_scene_create.restype = C.c_int
    lib.surf_scene_create.argtypes = [P(SceneInputs), vp, P(vp)]
    lib.surf_scene_destroy.restype = None
    lib.surf_scene_destroy.argtypes = [vp]
    lib.surf_scene_get_stats.restype = C.c_int
    lib.surf_scene_get_stats.argtypes = [vp, P(SceneStats)]
    lib.surf_scene_update_volume.restype = C.c_int
    lib.surf_scene_update_volume.argtypes = [vp, i32, vp, i64, vp]
    lib.surf_scene_set_views.restype = C.c_int
    lib.surf_scene_set_views.argtypes = [vp, P(SceneViews), vp]
    lib.surf_net_create.restype = C.c_int
    lib.surf_net_create.argtypes = [P(NetInputs), vp, P(vp)]
    lib.surf_net_destroy.restype = None
    lib.surf_net_destroy.argtypes = [vp]
    lib.surf_render_workspace_bytes.restype = C.c_size_t
    lib.surf_render_workspace_bytes.argtypes = [i64, i32, i32]
    lib.surf_sample_rays.restype = C.c_int
    lib.surf_sample_rays.argtypes = [vp, P(RenderCfg), vp, vp, vp, vp, vp, i64, vp, vp, vp]
    lib.surf_render_core.restype = C.c_int
    lib.surf_render_core.argtypes = [vp, vp, P(RenderCfg), vp, vp, vp, i64, i32, P(RenderOutputs), vp, C.c_size_t, vp]
    lib.surf_render_rays.restype = C.c_int
    lib.surf_render_rays.argtypes = [vp, vp, P(RenderCfg), vp, vp, vp, vp, vp, i64, P(RenderOutputs), vp,
                                     C.c_size_t, vp]
    lib.surf_sdf_points.restype = C.c_int
    lib.surf_sdf_points.argtypes = [vp, vp, vp, i64, vp, vp, i32, vp]
    lib.surf_sdf_full.restype = C.c_int
    lib.surf_sdf_full.argtypes = [vp, vp, vp, i64, vp, i32, vp]
    lib.surf_fpn_conv3x3.restype = C.c_int
    lib.surf_fpn_conv3x3.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.surf_fpn_deconv3x3s2.restype = C.c_int
    lib.surf_fpn_deconv3x3s2.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.surf_fpn_finish_stats.restype = C.c_int
    lib.surf_fpn_finish_stats.argtypes = [vp, i32, i32, i64, f32, vp, vp]
    lib.surf_fpn_conv_blocks.restype = C.c_int32
    lib.surf_fpn_conv_blocks.argtypes = [i32, i32]
    lib.surf_fpn_deconv_blocks.restype = C.c_int32
    lib.surf_fpn_deconv_blocks.argtypes = [i32, i32]
    lib.surf_fpn_norm_relu_add.restype = C.c_int
    lib.surf_fpn_norm_relu_add.argtypes = [vp, vp, vp, vp, i32, i64, vp, vp]
    lib.surf_mt19937_skip.restype = C.c_int
    lib.surf_mt19937_skip.argtypes = [vp, vp, vp, C.c_uint64]
    lib.surf_mask_dilate.restype = C.c_int
    lib.surf_mask_dilate.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    lib.surf_mesh_vertex_visibility.restype = C.c_int
    lib.surf_mesh_vertex_visibility.argtypes = [vp, i64, vp, vp, i32, vp, i32, i32, vp, vp]
    lib.surf_mesh_raster_workspace_bytes.restype = C.c_size_t
    lib.surf_mesh_raster_workspace_bytes.argtypes = [i32, i32]
    lib.surf_mesh_first_hits.restype = C.c_int
    lib.surf_mesh_first_hits.argtypes = [vp, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, vp, C.c_size_t, vp, vp, vp]
    lib.surf_mesh_components_workspace_bytes.restype = C.c_size_t
    lib.surf_mesh_components_workspace_bytes.argtypes = [i64]
    lib.surf_mesh_components.restype = C.c_int
    lib.surf_mesh_components.argtypes = [vp, i64, i32, vp, C.c_size_t, vp, vp, vp]
    lib.surf_sdf_smooth.restype = C.c_int
    lib.surf_sdf_smooth.argtypes = [vp, vp, vp, i64, vp, vp, vp, i32, vp]
    lib.surf_sdf_grid.restype = C.c_int
    lib.surf_sdf_grid.argtypes = [vp, vp, vp, i32, vp, i32, vp, i32, vp, i32, f32, i32, vp]
    lib.surf_point_mask.restype = C.c_int
    lib.surf_point_mask.argtypes = [vp, vp, i64, vp, vp]
    lib.surf_lookup_sparse.restype = C.c_int
    lib.surf_lookup_sparse.argtypes = [vp, vp, i64, vp, vp]
    lib.surf_lookup_feature.restype = C.c_int
    lib.surf_lookup_feature.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.surf_blend.restype = C.c_int
    lib.surf_blend.argtypes = [vp, vp, vp, vp, i64, i32, vp, i32, vp]
    lib.surf_tc_selftest.restype = C.c_int
    lib.surf_tc_selftest.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.surf_extras_workspace_bytes.restype = C.c_size_t
    lib.surf_extras_workspace_bytes.argtypes = [i64]
    lib.surf_render_extras.restype = C.c_int
    lib.surf_render_extras.argtypes = [vp, vp, P(ExtrasParams), vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, C.c_size_t, i32, vp]
    lib.surf_volume_back_proj.restype = C.c_int
    lib.surf_volume_back_proj.argtypes = [P(VolumeViews), P(vp), P(i32), P(i32), i32, i32, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.surf_volume_depth_filter.restype = C.c_int
    lib.surf_volume_depth_filter.argtypes = [P(VolumeViews), vp, vp, i64, vp, vp, f32, vp, vp]
    lib.surf_volume_upsample2x.restype = C.c_int
    lib.surf_volume_upsample2x.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.surf_scene_create_sparse.restype = C.c_int
    lib.surf_scene_create_sparse.argtypes = [P(SceneSparseInputs), vp, P(vp)]
    lib.surf_depth_map.restype = C.c_int
    lib.surf_depth_map.argtypes = [vp, P(DepthMapParams), vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.surf_mc_workspace_bytes.restype = C.c_size_t
    lib.surf_mc_workspace_bytes.argtypes = [i32, i32, i32]
    lib.surf_mc_count.restype = C.c_int
    lib.surf_mc_count.argtypes = [vp, i32, i32, i32, f32, vp, C.c_size_t, vp, vp]
    lib.surf_mc_emit.restype = C.c_int
    lib.surf_mc_emit.argtypes = [vp, i32, i32, i32, f32, vp, i32, vp, i64, vp, i64, vp]
    lib.surf_point_flags.restype = C.c_int
    lib.surf_point_flags.argtypes = [vp, P(RenderCfg), vp, vp, vp, i64, i32, vp, vp, vp, C.c_size_t, vp]


def load():
    """Loads the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "surf_b200: %s is missing — build it with `python -m surf_b200.build` "
                "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        missing = [n for n in EXPORTS if not hasattr(lib, n)]
        if missing:
            raise RuntimeError("surf_b200: library lacks symbols: %s" % ", ".join(missing))
        _declare(lib)
        if lib.surf_version() != ABI_VERSION:
            raise RuntimeError("surf_b200: ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().surf_last_error()
        raise RuntimeError("surf_b200.%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().surf_launch_count())


TIMING_KINDS = ["sdf_mlp_grad", "sdf_mlp_fwd", "lookup_feature", "blend", "sample_rays", "point_flags", "composite"]


def timing_enable(on=True):
    check(load().surf_timing_enable(1 if on else 0), "timing_enable")


def timing_read():
    """-> {kind: (ms_total, launches)} since the last read (synchronises the recorded events)."""
    ms = (C.c_double * len(TIMING_KINDS))()
    n = (C.c_int64 * len(TIMING_KINDS))()
    check(load().surf_timing_read(ms, n), "timing_read")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(TIMING_KINDS)}
