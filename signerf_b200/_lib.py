"""ctypes binding of the C ABI in include/signerf_b200.h.

This is the stub a SIGNeRF maintainer would add on the reference side: the
reference has no FFI of its own, its hot path is Python calling nerfstudio.  The library is
built in-tree by `__graft_entry__.build()` / `make -C signerf_b200/csrc`; there is NO CPU
fallback — a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsignerf_b200.so")

SGN_OK = 0
SGN_MLP_FP16_MMA = 0
SGN_MLP_FP32 = 1


class SgnError(RuntimeError):
    """A C-ABI call returned a negative SgnStatus."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"signerf_b200 error {code}: {msg}")
        self.code = code


class SgnHashGrid(C.Structure):
    _fields_ = [("d_table", C.c_void_p), ("h_scalings", C.POINTER(C.c_float)),
                ("num_levels", C.c_int), ("log2_size", C.c_int)]


class SgnLinear(C.Structure):
    _fields_ = [("h_weight", C.POINTER(C.c_float)), ("h_bias", C.POINTER(C.c_float)),
                ("in_dim", C.c_int), ("out_dim", C.c_int)]


class SgnFieldDesc(C.Structure):
    _fields_ = [("grid", SgnHashGrid), ("base", SgnLinear * 2), ("head", SgnLinear * 3),
                ("h_appearance", C.POINTER(C.c_float)), ("average_init_density", C.c_float),
                ("num_proposals", C.c_int), ("prop_grid", SgnHashGrid * 2),
                ("prop_mlp", (SgnLinear * 2) * 2)]


class SgnRenderOpts(C.Structure):
    _fields_ = [("near_plane", C.c_float), ("far_plane", C.c_float), ("mode", C.c_int),
                ("num_samples", C.c_int), ("num_prop_samples", C.c_int * 2), ("mlp_mode", C.c_int),
                ("h_bins", C.POINTER(C.c_float))]


class SgnMaskOpts(C.Structure):
    _fields_ = [("aabb", C.c_float * 6), ("inverse_mask", C.c_int), ("dilate_w", C.c_int),
                ("dilate_h", C.c_int), ("depth_radius", C.c_float), ("use_manual_depth", C.c_int),
                ("manual_min", C.c_float), ("manual_max", C.c_float)]


class SgnEpilogue(C.Structure):
    _fields_ = [("d_bias", C.c_void_p), ("d_rowbias", C.c_void_p), ("rows_per_batch", C.c_int),
                ("d_residual", C.c_void_p), ("ldo", C.c_int64), ("out_f16", C.c_int), ("geglu", C.c_int),
                ("nchw", C.c_int), ("act_silu", C.c_int)]


class SgnTrainSamples(C.Structure):
    _fields_ = [("d_spacing", C.c_void_p * 3), ("d_euclid", C.c_void_p * 3), ("d_sigma", C.c_void_p * 2),
                ("d_weights", C.c_void_p * 2)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/signerf_b200.h declares.
SIGNATURES = {
    "sgn_last_error": (C.c_char_p, []),
    "sgn_abi_version": (_i, []),
    "sgn_launch_count": (C.c_uint64, []),
    "sgn_set_option": (_i, [C.c_char_p, _i]),
    "sgn_field_create": (_i, [C.POINTER(SgnFieldDesc), C.POINTER(_vp)]),
    "sgn_field_destroy": (None, [_vp]),
    "sgn_render_views": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(SgnRenderOpts), _vp, _vp, _vp, _vp]),
    "sgn_attention_causal_f16": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp, _i64, _vp]),
    "sgn_act_f16": (_i, [_vp, _i64, _i, _vp, _vp]),
    "sgn_embed_tokens": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_mlp_param_count": (_i64, []),
    "sgn_field_mlp_params": (_i, [_vp, C.POINTER(_vp)]),
    "sgn_field_refresh": (_i, [_vp, _vp]),
    "sgn_train_forward": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgn_train_ws_bytes": (_i64, [_i64, _i]),
    "sgn_train_backward": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sgn_pred_normals_param_count": (_i64, []),
    "sgn_train_normals_forward": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp]),
    "sgn_normal_losses": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "sgn_train_normals_ws_bytes": (_i64, [_i64, _i]),
    "sgn_train_normals_backward": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sgn_train_normals_step": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sgn_appearance_bias": (_i, [_vp, _vp, _vp, _i, _vp, _i64, _vp, _vp]),
    "sgn_appearance_bias_backward": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "sgn_prop_param_count": (_i64, []),
    "sgn_field_prop_params": (_i, [_vp, _i, C.POINTER(_vp)]),
    "sgn_train_sample_ws_bytes": (_i64, [_i64, _i, _i]),
    "sgn_train_sample": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _f, _f, _vp, _f, C.POINTER(SgnTrainSamples), _vp, _i64, _vp]),
    "sgn_weights_from_density": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "sgn_interlevel_loss": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _f, _vp, _vp, _vp, _i64, _vp]),
    "sgn_distortion_loss": (_i, [_vp, _vp, _i64, _i, _f, _vp, _vp, _vp]),
    "sgn_prop_backward": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sgn_rgb_loss": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "sgn_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _vp]),
    "sgn_render_rays": (_i, [_vp, _vp, _vp, _i64, C.POINTER(SgnRenderOpts), _vp, _vp, _vp, _vp]),
    "sgn_render_views_host": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(SgnRenderOpts), _vp, _vp, _vp]),
    "sgn_generate_rays": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sgn_hash_encode": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _vp]),
    "sgn_field_eval": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "sgn_mask_condition": (_i, [_vp, _vp, _i, _i, _i, _vp, C.POINTER(SgnMaskOpts), _vp, _vp, _vp, _vp]),
    "sgn_mask_condition_combined": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, C.POINTER(SgnMaskOpts), _vp, _vp, _vp, _vp]),
    "sgn_dilate_ellipse": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_sheet_paste": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "sgn_sheet_cut": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "sgn_blend_masked": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "sgn_scatter_tiles_peer": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_quantize_u8": (_i, [_vp, _i64, _vp, _vp]),
    "sgn_gemm_f16": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, C.POINTER(SgnEpilogue), _vp, _vp]),
    "sgn_gemm_plan": (_i, [_i, _i, _i, _i, C.POINTER(_i)]),
    "sgn_conv3x3_f16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(SgnEpilogue), _vp, _vp]),
    "sgn_attention_f16": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _f, _vp, _i64, _vp]),
    "sgn_attention_f16_ws": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _f, _vp, _i64, _vp, _i64, _vp]),
    "sgn_attention_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "sgn_group_norm_ws_doubles": (_i64, [_i, _i, _i]),
    "sgn_group_norm_f16": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "sgn_layer_norm_f16": (_i, [_vp, _i64, _i, _f, _vp, _vp, _vp, _vp]),
    "sgn_cast_f16": (_i, [_vp, _i64, _vp, _vp]),
    "sgn_upsample2x_f16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_concat_f32": (_i, [_vp, _i, _vp, _vp, _f, _i, _i64, _vp, _vp]),
    "sgn_concat_f32_f16": (_i, [_vp, _i, _vp, _vp, _f, _i, _i64, _vp, _vp, _vp]),
    "sgn_axpy_f32": (_i, [_vp, _f, _i64, _vp, _vp]),
    "sgn_im2col3x3_s2_f16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_im2col3x3_split_f16": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_conv3x3_direct": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_linear_small": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_linear_small_segments": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "sgn_timestep_embedding": (_i, [_vp, _i, _i, _vp, _vp]),
    "sgn_scale_repeat_f32": (_i, [_vp, _i64, _f, _i, _vp, _vp]),
    "sgn_sheet_to_conditioning": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sgn_cfg_euler_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp]),
    "sgn_conv3x3_small_tc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sgn_im2col3x3_s2_asym_f16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_group_norm_split_f16": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "sgn_split_f16": (_i, [_vp, _i64, _i, _vp, _vp]),
    "sgn_upsample2x_split_f16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_im2col3x3_s2_asym_split_f16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sgn_softmax_rows_f16": (_i, [_vp, _i64, _i, _f, _vp, _vp]),
    "sgn_pointwise_nchw": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _i64, _f, _vp, _vp]),
    "sgn_vae_sample_latent": (_i, [_vp, _vp, _i, _i, _i64, _f, _vp, _vp]),
    "sgn_u8_to_vae_input": (_i, [_vp, _i, _i, _vp, _vp]),
    "sgn_vae_output_to_u8": (_i, [_vp, _i, _i, _vp, _vp]),
    "sgn_gaussian_blur_u8": (_i, [_vp, _i, _i, _i, C.c_double, _i, _vp, _vp]),
    "sgn_gaussian_kernel_q8": (_i, [_i, C.c_double, C.POINTER(_i)]),
    "sgn_pil_resize_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "sgn_pil_resize_bicubic_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sgn_inpaint_overlay_mask_u8": (_i, [_vp, _i64, _vp, _vp]),
    "sgn_latent_keep_mask": (_i, [_vp, _i64, _vp, _vp]),
    "sgn_overlay_composite_u8": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sgn_rasterize_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "sgn_rasterize_depth": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_double), _vp, _vp, _i, _i, _i, C.c_double, C.c_double, _i,
                                _vp, _vp, _vp]),
    "sgn_shape_color_u8": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "sgn_mask_condition_shape": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(SgnMaskOpts), _vp, _vp, _vp, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes.  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C signerf_b200/csrc`. signerf_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != SGN_OK:
        raise SgnError(code, load().sgn_last_error().decode("utf-8", "replace"))


def set_option(name: str, value: int) -> None:
    check(load().sgn_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return int(load().sgn_launch_count())
