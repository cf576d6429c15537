"""ctypes binding of libtod.so (include/tod.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libtod.so")

TOD_ACT_NONE, TOD_ACT_SILU = 0, 1
TOD_OUT_BF16, TOD_OUT_F32 = 0, 1

# every symbol include/tod.h declares (tests check the exports against the header)
SYMBOLS = (
    "tod_version", "tod_last_error", "tod_device_ok", "tod_conv2d_nhwc_bf16", "tod_conv_weight_layout",
    "tod_stem_conv_nchw_f32", "tod_sppf_pool_nhwc_bf16", "tod_head_decode", "tod_nms_prepare_dense",
    "tod_nms_workspace_bytes", "tod_nms", "tod_conv2d_nhwc_bf16_simt_check", "tod_decode_box_from_head",
    "tod_debug_set_conv_profile", "tod_stem_conv_nhwc_u8", "tod_conv2d_head_decode",
    "tod_resample_coeffs_bicubic", "tod_letterbox_bicubic_u8", "tod_correct_boxes", "tod_conv2d_tail1x1",
    "tod_conv2d_tail1x1_box_decode", "tod_cbam_workspace_floats", "tod_cbam_nhwc_bf16",
    "tod_softmax_rows_f32_bf16", "tod_attention_fused", "tod_transpose_bf16", "tod_decode_box_from_tuple",
    "tod_pack_workspace_bytes", "tod_pack_detections", "tod_debug_set_timeline", "tod_debug_timeline_launches",
    "tod_debug_timeline_name", "tod_set_sm_budget", "tod_get_sm_budget", "tod_loss_bbox_decode",
)


class ConvDesc(C.Structure):
    _fields_ = [
        ("d_x", C.c_void_p), ("d_w", C.c_void_p), ("d_bias", C.c_void_p), ("d_residual", C.c_void_p),
        ("d_upadd", C.c_void_p), ("d_out", C.c_void_p),
        ("batch", C.c_int32), ("hin", C.c_int32), ("win", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("x_pitch", C.c_int32), ("res_pitch", C.c_int32), ("out_pitch", C.c_int32),
        ("act", C.c_int32), ("out_dtype", C.c_int32), ("block_k", C.c_int32), ("num_stages", C.c_int32),
        ("reserved", C.c_int32 * 4), ("flags", C.c_int32), ("reserved2", C.c_int32 * 3),
    ]


TOD_CONV_DYNAMIC_W, TOD_CONV_REVERSE, TOD_CONV_PATCH_TILES = 1, 2, 4
TOD_CONV_PAIR_ON, TOD_CONV_PAIR_OFF = 8, 16


class DecodeDesc(C.Structure):
    _fields_ = [
        ("d_raw", C.c_void_p * 3), ("h", C.c_int32 * 3), ("w", C.c_int32 * 3), ("stride", C.c_float * 3),
        ("raw_pitch", C.c_int32), ("batch", C.c_int32), ("nc", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("d_head_out", C.c_void_p), ("d_decoded", C.c_void_p), ("d_cand_box", C.c_void_p),
        ("d_cand_conf", C.c_void_p), ("d_cand_cls", C.c_void_p),
        ("reserved", C.c_int32 * 4),
    ]


class HeadFuseDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("nc", C.c_int32), ("level_off", C.c_int32), ("anchors", C.c_int32),
        ("in_h", C.c_int32), ("in_w", C.c_int32), ("stride", C.c_float),
        ("d_cand_box", C.c_void_p), ("d_cand_conf", C.c_void_p), ("d_cand_cls", C.c_void_p),
        ("reserved", C.c_int32 * 4),
    ]


class LetterboxDesc(C.Structure):
    _fields_ = [
        ("d_src", C.c_void_p), ("d_tmp", C.c_void_p), ("d_dst", C.c_void_p),
        ("src_image_stride", C.c_int64), ("dst_image_stride", C.c_int64),
        ("n", C.c_int32), ("src_h", C.c_int32), ("src_w", C.c_int32), ("dst_h", C.c_int32), ("dst_w", C.c_int32),
        ("new_h", C.c_int32), ("new_w", C.c_int32), ("off_y", C.c_int32), ("off_x", C.c_int32), ("pad_value", C.c_int32),
        ("d_xbounds", C.c_void_p), ("d_xcoef", C.c_void_p), ("d_ybounds", C.c_void_p), ("d_ycoef", C.c_void_p),
        ("xksize", C.c_int32), ("yksize", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class CbamDesc(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("d_out", C.c_void_p), ("d_fc1", C.c_void_p), ("d_fc2", C.c_void_p),
                ("d_conv", C.c_void_p), ("d_work", C.c_void_p),
                ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("hidden", C.c_int32),
                ("ksize", C.c_int32), ("x_pitch", C.c_int32), ("out_pitch", C.c_int32), ("reserved", C.c_int32 * 4)]


class AttentionDesc(C.Structure):
    _fields_ = [("d_q", C.c_void_p), ("d_k", C.c_void_p), ("d_vt", C.c_void_p), ("d_bias", C.c_void_p), ("d_x", C.c_void_p),
                ("d_out", C.c_void_p), ("batch", C.c_int32), ("n", C.c_int32), ("c", C.c_int32), ("d16", C.c_int32),
                ("x_pitch", C.c_int32), ("out_pitch", C.c_int32), ("reserved", C.c_int32 * 4)]


class ConvTailDesc(C.Structure):
    _fields_ = [("d_w2", C.c_void_p), ("d_bias2", C.c_void_p), ("d_out2", C.c_void_p),
                ("cout2", C.c_int32), ("out2_pitch", C.c_int32), ("act2", C.c_int32), ("reserved", C.c_int32 * 5)]


TOD_FUSE_BOX, TOD_FUSE_CLS = 1, 2


class TodError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libtod.so (building it with nvcc if the in-tree binary is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from .build import build_library
        build_library()
    L = C.CDLL(LIB_PATH)
    L.tod_version.restype = C.c_int
    L.tod_last_error.restype = C.c_char_p
    L.tod_device_ok.restype = C.c_int
    L.tod_conv2d_nhwc_bf16.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    L.tod_conv2d_nhwc_bf16_simt_check.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    L.tod_conv2d_head_decode.argtypes = [C.POINTER(ConvDesc), C.POINTER(HeadFuseDesc), C.c_void_p]
    L.tod_conv_weight_layout.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]
    L.tod_stem_conv_nchw_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_void_p]
    L.tod_stem_conv_nhwc_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]
    L.tod_sppf_pool_nhwc_bf16.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.tod_head_decode.argtypes = [C.POINTER(DecodeDesc), C.c_void_p]
    L.tod_nms_prepare_dense.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    L.tod_nms_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.tod_nms_workspace_bytes.restype = C.c_int64
    L.tod_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_double, C.c_void_p,
                          C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tod_decode_box_from_head.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p]
    L.tod_decode_box_from_tuple.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.tod_pack_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.tod_pack_workspace_bytes.restype = C.c_int64
    L.tod_pack_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]
    L.tod_debug_set_conv_profile.argtypes = [C.c_void_p]
    L.tod_set_sm_budget.argtypes = [C.c_int32]
    L.tod_loss_bbox_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.tod_debug_set_timeline.argtypes = [C.c_void_p]
    L.tod_debug_timeline_name.argtypes = [C.c_int]
    L.tod_debug_timeline_name.restype = C.c_char_p
    L.tod_resample_coeffs_bicubic.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.tod_letterbox_bicubic_u8.argtypes = [C.POINTER(LetterboxDesc), C.c_void_p]
    L.tod_conv2d_tail1x1.argtypes = [C.POINTER(ConvDesc), C.POINTER(ConvTailDesc), C.c_void_p]
    L.tod_conv2d_tail1x1_box_decode.argtypes = [C.POINTER(ConvDesc), C.POINTER(ConvTailDesc), C.POINTER(HeadFuseDesc), C.c_void_p]
    L.tod_cbam_workspace_floats.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.tod_cbam_workspace_floats.restype = C.c_int64
    L.tod_cbam_nhwc_bf16.argtypes = [C.POINTER(CbamDesc), C.c_void_p]
    L.tod_softmax_rows_f32_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    L.tod_transpose_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    L.tod_attention_fused.argtypes = [C.POINTER(AttentionDesc), C.c_void_p]
    L.tod_correct_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    for name in SYMBOLS:
        getattr(L, name)  # fail loudly if the binary is stale
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise TodError(f"{what} failed ({rc}): {lib().tod_last_error().decode(errors='replace')}")


def weight_layout(cin: int, ksize: int, block_k_hint: int = 0):
    bk, cp, kt = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().tod_conv_weight_layout(cin, ksize, block_k_hint, C.byref(bk), C.byref(cp), C.byref(kt)), "tod_conv_weight_layout")
    return bk.value, cp.value, kt.value
