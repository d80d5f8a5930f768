"""ctypes binding of libd2gs.so (include/d2gs.h).

The product path has NO fallback: if the CUDA library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# D2GS_LIB: another build of the same C ABI (A/B timing of kernel variants); default = the in-tree library
LIB_PATH = os.environ.get("D2GS_LIB") or os.path.join(_HERE, "libd2gs.so")

c_float_p = C.c_void_p  # raw device addresses are passed as integers


class D2gsConfig(C.Structure):
    _fields_ = [("num_channels", C.c_int), ("block_x", C.c_int), ("block_y", C.c_int), ("tight_bbox", C.c_int),
                ("render_auxiliary", C.c_int), ("backface_cull", C.c_int), ("dual_visible", C.c_int),
                ("detach_weight", C.c_int), ("near_plane", C.c_double), ("far_plane", C.c_double),
                ("filter_size", C.c_double), ("sm_arch", C.c_int)]


class RasterFwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("sh_rest", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
                ("rotations", C.c_void_p), ("transMat_precomp", C.c_void_p), ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("prefiltered", C.c_int), ("debug", C.c_int),
                ("out_color", C.c_void_p), ("out_others", C.c_void_p), ("radii", C.c_void_p),
                ("geom_buffer", C.c_void_p), ("geom_bytes", C.c_size_t),
                ("img_buffer", C.c_void_p), ("img_bytes", C.c_size_t),
                ("binning_buffer", C.c_void_p), ("binning_bytes", C.c_size_t),
                ("resume", C.c_int),
                ("num_rendered", C.POINTER(C.c_int64)), ("binning_required", C.POINTER(C.c_size_t)),
                ("raw_params", C.c_int), ("d_means3D", C.c_void_p), ("d_scales", C.c_void_p), ("d_rotations", C.c_void_p),
                ("binning_capacity", C.c_int64), ("num_rendered_async", C.c_void_p)]


class Gs3dFwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
                ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
                ("scale_modifier", C.c_float), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("prefiltered", C.c_int), ("debug", C.c_int),
                ("out_color", C.c_void_p), ("out_depth", C.c_void_p), ("out_alpha", C.c_void_p), ("radii", C.c_void_p),
                ("geom_buffer", C.c_void_p), ("geom_bytes", C.c_size_t), ("img_buffer", C.c_void_p), ("img_bytes", C.c_size_t),
                ("binning_buffer", C.c_void_p), ("binning_bytes", C.c_size_t), ("resume", C.c_int),
                ("num_rendered", C.POINTER(C.c_int64)), ("binning_required", C.POINTER(C.c_size_t)),
                ("binning_capacity", C.c_int64), ("num_rendered_async", C.c_void_p)]


class Gs3dBwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("num_rendered", C.c_int64),
                ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
                ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("radii", C.c_void_p), ("out_alpha", C.c_void_p),
                ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("img_buffer", C.c_void_p),
                ("dL_dout_color", C.c_void_p), ("dL_dout_depth", C.c_void_p), ("dL_dout_alpha", C.c_void_p),
                ("debug", C.c_int), ("grad_scratch", C.c_void_p),
                ("dL_dmeans2D", C.c_void_p), ("dL_dcolors", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dmeans3D", C.c_void_p),
                ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p), ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p)]


class Gs3dState(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("cov3D", C.c_void_p), ("clamped", C.c_void_p), ("tiles_touched", C.c_void_p),
                ("keys_sorted", C.c_void_p), ("point_list", C.c_void_p), ("ranges", C.c_void_p), ("n_contrib", C.c_void_p)]


class RasterBwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("num_rendered", C.c_int64),
                ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("sh_rest", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("transMat_precomp", C.c_void_p), ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("radii", C.c_void_p), ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p),
                ("img_buffer", C.c_void_p), ("dL_dout_color", C.c_void_p), ("dL_dout_others", C.c_void_p),
                ("debug", C.c_int), ("grad_scratch", C.c_void_p),
                ("dL_dmeans2D", C.c_void_p), ("dL_dcolors", C.c_void_p), ("dL_dopacity", C.c_void_p),
                ("dL_dmeans3D", C.c_void_p), ("dL_dtransMat", C.c_void_p), ("dL_dsh", C.c_void_p),
                ("dL_dsh_rest", C.c_void_p), ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
                ("raw_params", C.c_int), ("opacities", C.c_void_p), ("d_means3D", C.c_void_p), ("d_scales", C.c_void_p),
                ("d_rotations", C.c_void_p), ("dL_dscales_raw", C.c_void_p)]


class RasterState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("means2D", "depths", "transMat", "normal_opacity", "rgb", "clamped", "tiles_touched", "point_offsets",
                 "keys_unsorted", "values_unsorted", "keys_sorted", "point_list", "ranges", "final_T", "n_contrib")]


class DeformFwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("M", C.c_int), ("K", C.c_int), ("hyper_dim", C.c_int),
                ("xyz", C.c_void_p), ("feature", C.c_void_p), ("feature_stride", C.c_int),
                ("nodes", C.c_void_p), ("node_radius_log", C.c_void_p), ("node_weight_logit", C.c_void_p),
                ("node_trans", C.c_void_p), ("node_rot", C.c_void_p), ("node_scale", C.c_void_p),
                ("node_local_rot", C.c_void_p), ("motion_mask", C.c_void_p),
                ("nn_idx", C.c_void_p), ("nn_dist", C.c_void_p), ("nn_weight", C.c_void_p),
                ("d_xyz", C.c_void_p), ("d_rotation", C.c_void_p), ("d_scaling", C.c_void_p), ("node_attr_stride", C.c_int),
                ("order", C.c_void_p)]


class DeformBwdArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("M", C.c_int), ("K", C.c_int), ("hyper_dim", C.c_int),
                ("xyz", C.c_void_p), ("feature", C.c_void_p), ("feature_stride", C.c_int),
                ("nodes", C.c_void_p), ("node_radius_log", C.c_void_p), ("node_weight_logit", C.c_void_p),
                ("node_trans", C.c_void_p), ("node_rot", C.c_void_p), ("node_scale", C.c_void_p),
                ("node_local_rot", C.c_void_p), ("motion_mask", C.c_void_p),
                ("nn_idx", C.c_void_p), ("nn_dist", C.c_void_p), ("nn_weight", C.c_void_p),
                ("dL_d_xyz", C.c_void_p), ("dL_d_rotation", C.c_void_p), ("dL_d_scaling", C.c_void_p),
                ("dL_dnode_trans", C.c_void_p), ("dL_dnode_rot", C.c_void_p), ("dL_dnode_scale", C.c_void_p),
                ("dL_dnode_local_rot", C.c_void_p), ("dL_dnodes", C.c_void_p), ("dL_dnode_radius_log", C.c_void_p),
                ("dL_dnode_weight_logit", C.c_void_p), ("dL_dfeature", C.c_void_p), ("dL_dmotion_mask", C.c_void_p),
                ("node_attr_stride", C.c_int), ("order", C.c_void_p)]


class EpilogueArgs(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("allmap", C.c_void_p), ("viewmatrix", C.c_void_p),
                ("focal_x", C.c_float), ("focal_y", C.c_float),
                ("alpha", C.c_void_p), ("rend_normal", C.c_void_p), ("rend_dist", C.c_void_p), ("depth", C.c_void_p),
                ("surf_normal", C.c_void_p), ("surf_point", C.c_void_p),
                ("g_alpha", C.c_void_p), ("g_rend_normal", C.c_void_p), ("g_rend_dist", C.c_void_p), ("g_depth", C.c_void_p),
                ("g_surf_normal", C.c_void_p), ("g_surf_point", C.c_void_p), ("dL_dallmap", C.c_void_p)]


class LossArgs(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("image", C.c_void_p), ("gt", C.c_void_p),
                ("rend_normal", C.c_void_p), ("surf_normal", C.c_void_p), ("rend_dist", C.c_void_p),
                ("lambda_dssim", C.c_float), ("lambda_normal", C.c_float), ("lambda_dist", C.c_float),
                ("out", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("save_for_backward", C.c_int),
                ("upstream", C.c_void_p), ("g_image", C.c_void_p), ("g_rend_normal", C.c_void_p), ("g_surf_normal", C.c_void_p),
                ("g_rend_dist", C.c_void_p)]


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_float),
                ("step_size", C.c_float), ("bias_correction2_sqrt", C.c_float)]


class MlpArgs(C.Structure):
    _fields_ = [("rows", C.c_int), ("is_blender", C.c_int), ("num_out", C.c_int), ("x", C.c_void_p), ("t", C.c_void_p),
                ("t_stride", C.c_int),
                ("timenet0_w", C.c_void_p), ("timenet0_b", C.c_void_p), ("timenet2_w", C.c_void_p), ("timenet2_b", C.c_void_p),
                ("linear_w", C.c_void_p * 8), ("linear_b", C.c_void_p * 8), ("heads_w", C.c_void_p), ("heads_b", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("out", C.c_void_p), ("g_out", C.c_void_p),
                ("g_timenet0_w", C.c_void_p), ("g_timenet0_b", C.c_void_p), ("g_timenet2_w", C.c_void_p), ("g_timenet2_b", C.c_void_p),
                ("g_linear_w", C.c_void_p * 8), ("g_linear_b", C.c_void_p * 8), ("g_heads_w", C.c_void_p), ("g_heads_b", C.c_void_p)]


# every symbol include/d2gs.h declares; tests assert the shared library exports all of them
EXPORTED_SYMBOLS = (
    "d2gs_set_option", "d2gs_profile_enable", "d2gs_profile_collect", "d2gs_last_error", "d2gs_version", "d2gs_get_config", "d2gs_raster_workspace", "d2gs_raster_forward",
    "d2gs_raster_backward", "d2gs_mark_visible", "d2gs_raster_export_state", "d2gs_deform_forward",
    "d2gs_deform_backward", "d2gs_epilogue_forward", "d2gs_epilogue_backward",
    "d2gs_mlp_workspace", "d2gs_mlp_forward", "d2gs_mlp_backward", "d2gs_mlp_hidden",
    "d2gs_deform_order_workspace", "d2gs_deform_order", "d2gs_knn_mean_dist2_workspace", "d2gs_knn_mean_dist2",
    "d2gs_gs3d_workspace", "d2gs_gs3d_forward", "d2gs_gs3d_backward", "d2gs_gs3d_export_state",
    "d2gs_loss_workspace", "d2gs_loss_forward", "d2gs_loss_backward",
    "d2gs_adam_step", "d2gs_densification_stats",
)

D2GS_OK = 0
D2GS_NEED_BINNING = 1

_lib = None


class D2gsError(RuntimeError):
    pass


def lib():
    """Load libd2gs.so.  Raises (never falls back) when the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D2gsError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C dynamic-2dgs_b200/csrc`). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.d2gs_last_error.restype = C.c_char_p
    L.d2gs_version.restype = C.c_char_p
    L.d2gs_get_config.argtypes = [C.POINTER(D2gsConfig)]
    L.d2gs_raster_workspace.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.d2gs_raster_forward.argtypes = [C.POINTER(RasterFwdArgs), C.c_void_p]
    L.d2gs_raster_backward.argtypes = [C.POINTER(RasterBwdArgs), C.c_void_p]
    L.d2gs_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.d2gs_raster_export_state.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(RasterState), C.c_void_p]
    L.d2gs_deform_forward.argtypes = [C.POINTER(DeformFwdArgs), C.c_void_p]
    L.d2gs_deform_backward.argtypes = [C.POINTER(DeformBwdArgs), C.c_void_p]
    L.d2gs_epilogue_forward.argtypes = [C.POINTER(EpilogueArgs), C.c_void_p]
    L.d2gs_epilogue_backward.argtypes = [C.POINTER(EpilogueArgs), C.c_void_p]
    L.d2gs_mlp_workspace.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.d2gs_mlp_forward.argtypes = [C.POINTER(MlpArgs), C.c_void_p]
    L.d2gs_mlp_backward.argtypes = [C.POINTER(MlpArgs), C.c_void_p]
    L.d2gs_mlp_hidden.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.d2gs_mlp_hidden.restype = C.c_void_p
    L.d2gs_deform_order_workspace.argtypes = [C.c_int, C.POINTER(C.c_size_t)]
    L.d2gs_deform_order.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.d2gs_gs3d_workspace.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_size_t)]
    L.d2gs_gs3d_forward.argtypes = [C.POINTER(Gs3dFwdArgs), C.c_void_p]
    L.d2gs_gs3d_backward.argtypes = [C.POINTER(Gs3dBwdArgs), C.c_void_p]
    L.d2gs_gs3d_export_state.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(Gs3dState), C.c_void_p]
    L.d2gs_knn_mean_dist2_workspace.argtypes = [C.c_int, C.POINTER(C.c_size_t)]
    L.d2gs_knn_mean_dist2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.d2gs_loss_workspace.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.d2gs_loss_forward.argtypes = [C.POINTER(LossArgs), C.c_void_p]
    L.d2gs_loss_backward.argtypes = [C.POINTER(LossArgs), C.c_void_p]
    L.d2gs_adam_step.argtypes = [C.POINTER(AdamTensor), C.c_int, C.c_void_p]
    L.d2gs_densification_stats.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.d2gs_set_option.argtypes = [C.c_char_p, C.c_int]
    L.d2gs_profile_enable.argtypes = [C.c_int]
    L.d2gs_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for name in EXPORTED_SYMBOLS:
        getattr(L, name)
    _lib = L
    # D2GS_OPTIONS="tile_sort=1,cull=0": runtime switches for scripts that have no flag of their own (profiling targets)
    for item in filter(None, os.environ.get("D2GS_OPTIONS", "").split(",")):
        k, _, v = item.partition("=")
        check(L.d2gs_set_option(k.strip().encode(), int(v or 1)), "d2gs_set_option")
    return L


def check(code: int, what: str) -> None:
    if code < 0:
        raise D2gsError(f"{what} failed ({code}): {lib().d2gs_last_error().decode()}")


def config() -> dict:
    c = D2gsConfig()
    check(lib().d2gs_get_config(C.byref(c)), "d2gs_get_config")
    return {n: getattr(c, n) for n, _ in D2gsConfig._fields_}


def workspace_sizes(P: int, W: int, H: int, R: int = 0):
    g, i, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
    check(lib().d2gs_raster_workspace(P, W, H, R, C.byref(g), C.byref(i), C.byref(b)), "d2gs_raster_workspace")
    return g.value, i.value, b.value


STAGE_NAMES = ("preprocess_fwd", "scan", "duplicate", "sort", "ranges", "blend_fwd", "blend_bwd", "preprocess_bwd",
               "deform_fwd", "deform_bwd", "epilogue_fwd", "epilogue_bwd", "mlp_fwd", "mlp_bwd", "loss_fwd", "loss_bwd")


def profile_enable(on: bool) -> None:
    check(lib().d2gs_profile_enable(int(on)), "d2gs_profile_enable")


def profile_collect() -> dict:
    """{stage: (total_ms, launches)} since the last collect; synchronises the device."""
    n = len(STAGE_NAMES)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib().d2gs_profile_collect(ms, cnt), "d2gs_profile_collect")
    return {STAGE_NAMES[i]: (ms[i], cnt[i]) for i in range(n)}


def set_option(name: str, value: int) -> None:
    check(lib().d2gs_set_option(name.encode(), int(value)), "d2gs_set_option")
