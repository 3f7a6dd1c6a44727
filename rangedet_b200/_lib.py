"""ctypes binding of librangedet_b200.so (the C-ABI in include/rangedet_b200.h).

There is no CPU or eager-PyTorch fallback: if the library is missing, or a compute entry point
reports an error, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librangedet_b200.so")

_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t
_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/rangedet_b200.h one to one
SIGNATURES = {
    "rd_version": (_i, []),
    "rd_last_error": (ctypes.c_char_p, []),
    "rd_check_device": (_i, []),
    "rd_launch_count": (ctypes.c_uint64, []),
    "rd_set_pdl": (_i, [_i]),
    "rd_set_conv_t": (_i, [_i]),
    "rd_meta_kernel_fwd": (_i, [_vp] * 7 + [_i] * 5 + [_vp]),
    "rd_meta_kernel_fwd_nhwc_bf16": (_i, [_vp] * 8 + [_i, _vp] + [_i] * 4 + [_vp]),
    "rd_meta_kernel_bwd_workspace_bytes": (_sz, [_i] * 4),
    "rd_meta_kernel_bwd": (_i, [_vp] * 12 + [_vp, _sz] + [_i] * 5 + [_vp]),
    "rd_meta_kernel_bwd_nhwc_bf16": (_i, [_vp] * 12 + [_vp, _sz] + [_i] * 4 + [_vp]),
    "rd_meta_kernel_bwd_nhwc_f16": (_i, [_vp] * 12 + [_vp, _sz] + [_i] * 4 + [_vp]),
    "rd_meta_kernel_bwd_data": (_i, [_vp] * 7 + [_i] * 5 + [_vp]),
    "rd_meta_kernel_bwd_params": (_i, [_vp] * 11 + [_vp, _sz] + [_i] * 5 + [_vp]),
    "rd_decode_3d_bbox": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "rd_rotated_iou": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _vp]),
    "rd_batch_rotated_iou_max": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp]),
    "rd_rpn_loss_workspace_bytes": (_sz, []),
    "rd_rpn_loss": (_i, [_vp] * 8 + [_i, _i64, _i, _i] + [_f] * 6 + [_vp] * 5 + [_vp, _sz, _vp]),
    "rd_rpn_loss_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i] + [_vp] * 6 + [_i, _i, _i] + [_f] * 6 + [_vp] * 5 + [_vp, _sz, _vp]),
    "rd_rpn_loss_nhwc_f16": (_i, [_vp, _vp, _i, _i, _i] + [_vp] * 6 + [_i, _i, _i] + [_f] * 6 + [_vp] * 5 + [_vp, _sz, _vp]),
    "rd_assign3d_v2": (_i, [_vp] * 6 + [_f] * 7 + [_i64, _i, _vp, _vp]),
    "rd_get_point_num_workspace_bytes": (_sz, []),
    "rd_get_point_num": (_i, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "rd_rpn_reg_target": (_i, [_vp] * 5 + [_i64, _i] + [_vp] * 3 + [_vp]),
    "rd_wnms_4c_workspace_bytes": (_sz, [_i]),
    "rd_wnms_4c": (_i, [_vp, _i, _f, _f, _i, _i, _vp, _vp, ctypes.POINTER(_i), _vp, _sz, _vp]),
    "rd_nms3d_workspace_bytes": (_sz, [_i, _i]),
    "rd_nms3d": (_i, [_vp, _i, _i, _f, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "rd_get_sorted_foreground_workspace_bytes": (_sz, [_i, _i]),
    "rd_get_sorted_foreground": (_i, [_vp] * 4 + [_i] * 3 + [_vp] * 4 + [_sz, _vp]),
    "rd_conv2d_nhwc_bf16": (_i, [_vp] * 6 + [_i] * 8 + [_vp]),
    "rd_conv2d_nhwc_bf16_slice": (_i, [_vp] * 5 + [_i] * 10 + [_vp]),
    "rd_conv2d_nhwc_bf16_stats": (_i, [_vp] * 3 + [_i] * 7 + [_vp, _sz, ctypes.POINTER(_i), _vp]),
    "rd_conv2d_nhwc_bf16_bwdstats": (_i, [_vp] * 5 + [_i] * 6 + [_vp, _sz, ctypes.POINTER(_i), _vp]),
    "rd_bn_act_bwd_apply_nhwc_bf16": (_i, [_vp] * 4 + [_i, _vp, _i, _vp, _i, _vp, _vp, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "rd_bn_train_finalize": (_i, [_vp] + [_i] * 5 + [_vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "rd_deconv2d_nhwc_bf16": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "rd_conv2d_wgrad_workspace_bytes": (_sz, [_i] * 7),
    "rd_conv2d_wgrad_nhwc_bf16": (_i, [_vp] * 3 + [_i] * 7 + [_vp, _sz, _vp]),
    "rd_bn_workspace_bytes": (_sz, [_i]),
    "rd_bn_train_stats_nhwc_bf16": (_i, [_vp] + [_i] * 4 + [_vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rd_bn_act_fwd_nhwc_bf16": (_i, [_vp] * 5 + [_i] * 5 + [_vp]),
    "rd_bn_act_bwd_nhwc_bf16": (_i, [_vp] * 4 + [_i, _vp, _i, _vp, _vp, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "rd_channel_sums_nhwc_bf16": (_i, [_vp] + [_i] * 4 + [_vp, _vp, _sz, _vp]),
    "rd_add_nhwc_bf16": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "rd_nhwc_bf16_to_nchw_f32": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "rd_nchw_f32_to_nhwc_bf16": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "rd_gather_f32_to_bf16": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "rd_gather_f32": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "rd_copy_channels_16b": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i64, _vp]),
    "rd_sgd_mom_update": (_i, [_vp] * 5 + [_i64, _vp]),
    # fp16-storage twins of every *bf16* entry point (same arguments)
    "rd_meta_kernel_fwd_nhwc_f16": (_i, [_vp] * 8 + [_i, _vp] + [_i] * 4 + [_vp]),
    "rd_conv2d_nhwc_f16": (_i, [_vp] * 6 + [_i] * 8 + [_vp]),
    "rd_conv2d_nhwc_f16_slice": (_i, [_vp] * 5 + [_i] * 10 + [_vp]),
    "rd_conv2d_nhwc_f16_stats": (_i, [_vp] * 3 + [_i] * 7 + [_vp, _sz, ctypes.POINTER(_i), _vp]),
    "rd_conv2d_nhwc_f16_bwdstats": (_i, [_vp] * 5 + [_i] * 6 + [_vp, _sz, ctypes.POINTER(_i), _vp]),
    "rd_bn_act_bwd_apply_nhwc_f16": (_i, [_vp] * 4 + [_i, _vp, _i, _vp, _i, _vp, _vp, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "rd_deconv2d_nhwc_f16": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "rd_conv2d_wgrad_nhwc_f16": (_i, [_vp] * 3 + [_i] * 7 + [_vp, _sz, _vp]),
    "rd_bn_train_stats_nhwc_f16": (_i, [_vp] + [_i] * 4 + [_vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rd_bn_act_fwd_nhwc_f16": (_i, [_vp] * 5 + [_i] * 5 + [_vp]),
    "rd_bn_act_bwd_nhwc_f16": (_i, [_vp] * 4 + [_i, _vp, _i, _vp, _vp, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "rd_channel_sums_nhwc_f16": (_i, [_vp] + [_i] * 4 + [_vp, _vp, _sz, _vp]),
    "rd_add_nhwc_f16": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "rd_nhwc_f16_to_nchw_f32": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "rd_nchw_f32_to_nhwc_f16": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "rd_gather_f32_to_f16": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "rd_tc_probe_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rd_tma_probe": (_i, [_vp, _vp, _vp] + [_i] * 7 + [_vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "rangedet_b200: %s not found. Build it with `python -m rangedet_b200.build` "
                "(there is no CPU fallback)." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def act_fn(name_bf16, dtype):
    """The entry point for the storage type of the tensors at hand: `name_bf16` (a name containing 'bf16') for
    torch.bfloat16, its fp16 twin for torch.float16."""
    import torch
    if dtype == torch.bfloat16:
        return getattr(lib(), name_bf16)
    if dtype == torch.float16:
        return getattr(lib(), name_bf16.replace("bf16", "f16"))
    raise TypeError("rangedet_b200: activations must be stored as bfloat16 or float16, got %s" % (dtype,))


def last_error():
    return lib().rd_last_error().decode("utf-8", "replace")


def check(status, what):
    if status != 0:
        raise RuntimeError("rangedet_b200.%s failed: %s" % (what, last_error()))


def set_pdl(on):
    """Programmatic dependent launch on / off for every subsequent kernel launch of this process (default on);
    returns the previous setting."""
    return bool(lib().rd_set_pdl(1 if on else 0))


def set_conv_t(on):
    """Transposed-orientation kernel (csrc/conv_t.cu) for 3x3 / stride-1 / Cout-128 convolutions on / off; returns the
    previous setting.  on = 160 / 192 / 224 / 256 also fixes the pixel-tile width (default: chosen per shape)."""
    return int(lib().rd_set_conv_t(int(on)))


def launch_count():
    return int(lib().rd_launch_count())
