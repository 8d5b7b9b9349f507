"""ctypes binding of libjg_b200.so (the C ABI declared in include/jg_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
Tensors are passed as raw device pointers; every call runs on torch's current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libjg_b200.so")

_lib = None

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_f = ctypes.c_float
c_p = ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    """Mirror of jg_conv_desc."""
    _fields_ = [
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("Cin", c_int), ("ldx", c_int),
        ("Ho", c_int), ("Wo", c_int),
        ("Cout", c_int), ("ldy", c_int),
        ("R", c_int), ("S", c_int),
        ("stride", c_int), ("pad", c_int), ("up2x", c_int), ("act", c_int),
        ("ldres", c_int), ("res_scale", c_f),
    ]


class ConvEpilogue(ctypes.Structure):
    """Mirror of jg_conv_epilogue (GroupNorm work fused into the convolution epilogue)."""
    _fields_ = [("stats", c_p), ("gn_sums", c_p), ("gn_x", c_p), ("ldgx", c_int), ("gn_ab", c_p), ("gn_act", c_int)]


class LinearItem(ctypes.Structure):  # jg_linear_item
    _fields_ = [("w", c_p), ("b", c_p), ("O", c_int), ("off", c_int)]


class PackItem(ctypes.Structure):  # jg_pack_item
    _fields_ = [("w", c_p), ("wf", c_p), ("wd", c_p), ("Cout", c_int), ("Cin", c_int), ("RS", c_int),
                ("Cin8", c_int), ("Cout8", c_int), ("pad_", c_int)]


class UnpackItem(ctypes.Structure):  # jg_unpack_item
    _fields_ = [("acc", c_p), ("dw", c_p), ("Cout", c_int), ("Cin", c_int), ("RS", c_int), ("layout", c_int)]


ACT_NONE, ACT_RELU, ACT_LRELU02, ACT_TANH, ACT_SILU = 0, 1, 2, 3, 4

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_SIGNATURES = {
    "jg_version": [],
    "jg_check_device": [],
    "jg_conv2d_fwd": [ctypes.POINTER(ConvDesc), c_p, c_p, c_p, c_p, c_p, c_p],
    "jg_conv2d_fwd_ex": [ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvEpilogue), c_p, c_p, c_p, c_p, c_p, c_p],
    "jg_chan_stats": [c_p, c_int, c_int, c_int, c_int, c_p, c_p],
    "jg_fill_mask_random": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "jg_u8_to_f32_normalized": [c_p, c_p, c_int, c_int, c_int, c_int, c_f, c_f, c_p],
    "jg_mask_class_dropout": [c_p, c_p, c_p, c_f, c_i64, c_p, c_p, c_int, c_i64, c_p],
    "jg_haar": [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_embed_rows": [c_p, c_p, c_p, c_p, c_int, c_int, c_i64, c_int, c_int, c_p],
    "jg_embed_rows_bwd": [c_p, c_int, c_int, c_p, c_p, c_i64, c_int, c_int, c_p, c_p, c_p],
    "jg_linear_batched_tiles": [c_int],
    "jg_linear_batched_fwd": [c_p, c_p, c_p, c_int, c_int, c_p, c_int, c_int, c_int, c_p],
    "jg_linear_batched_bwd": [c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "jg_comm_unique_id": [c_p],
    "jg_comm_init": [c_p, c_int, c_int, ctypes.POINTER(c_p)],
    "jg_comm_allreduce_async": [c_p, c_p, ctypes.c_size_t, c_int, c_p],
    "jg_comm_broadcast": [c_p, c_p, ctypes.c_size_t, c_int, c_p],
    "jg_comm_wait": [c_p, c_p],
    "jg_comm_info": [c_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_ulonglong),
                     ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(c_int)],
    "jg_comm_destroy": [c_p],
    "jg_conv2d_wgrad": [ctypes.POINTER(ConvDesc), c_p, c_p, c_int, c_p, c_p, c_f, c_p],
    "jg_pack_conv_weight": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "jg_conv2d_wgrad_acc": [ctypes.POINTER(ConvDesc), c_p, c_p, c_int, c_p, ctypes.POINTER(c_int), c_p],
    "jg_pack_conv_weights_batched": [c_p, c_p, c_int, c_int, c_p],
    "jg_wgrad_unpack_batched": [c_p, c_p, c_int, c_int, c_p],
    "jg_weight_tiles": [c_int, c_int, c_int],
    "jg_unpack_conv_wgrad": [c_p, c_p, c_int, c_int, c_int, c_int, c_f, c_p],
    "jg_bias_grad": [c_p, c_i64, c_int, c_int, c_p, c_p],
    "jg_nchw_f32_to_nhwc_bf16": [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_nhwc_bf16_to_nchw_f32": [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_copy_channels": [c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_p],
    "jg_resample2x": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_groupnorm_fwd": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_f, c_p, c_p, c_p, c_int, c_p, c_p,
                         c_p, c_p, c_p],
    "jg_groupnorm_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p,
                         c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "jg_attn_fwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_attn_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_layernorm_fwd": [c_p, c_int, c_p, c_int, c_i64, c_int, c_f, c_p, c_p, c_p, c_int, c_int, c_p, c_p],
    "jg_layernorm_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_p, c_p, c_p, c_p, c_p, c_p],
    "jg_temporal_attn_fwd": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_temporal_attn_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_ddpm_step": [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int,
                     c_p],
    "jg_rmsnorm_mod_fwd": [c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_f, c_p, c_p, c_p, c_int, c_p, c_p],
    "jg_rmsnorm_mod_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_p, c_p, c_int, c_p, c_p, c_p, c_p,
                           c_int, c_p],
    "jg_qknorm_rope_fwd": [c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_int, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "jg_qknorm_rope_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p,
                           c_p, c_p],
    "jg_attn_small_fwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_p],
    "jg_attn_small_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_p, c_int, c_p, c_int, c_p,
                          c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_swiglu_fwd": [c_p, c_int, c_p, c_int, c_i64, c_int, c_p],
    "jg_swiglu_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_p],
    "jg_gated_residual_fwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_p],
    "jg_gated_residual_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_p],
    "jg_geglu_fwd": [c_p, c_int, c_p, c_int, c_i64, c_int, c_p],
    "jg_geglu_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_p],
    "jg_linear_fwd": [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_linear_bwd": [c_p, c_p, c_p, c_p, c_int, c_p, c_p, c_int, c_int, c_int, c_int, c_p],
    "jg_noise_pack_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_palette_loss_fwd": [c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_int, c_p, c_p],
    "jg_palette_loss_bwd": [c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_int, c_p, c_p, c_int, c_p],
    "jg_adamw_ema_step": [c_p, c_p, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_f, c_int, c_int, c_p, c_f, c_f,
                          c_int, c_p],
    "jg_pad2d_fwd": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_pad2d_bwd": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_dilate2x": [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_act_bwd": [c_p, c_int, c_p, c_int, c_p, c_int, c_i64, c_int, c_int, c_p],
    "jg_gan_loss_fwd": [c_p, c_int, c_i64, c_int, c_int, c_f, c_f, c_p, c_p],
    "jg_gan_loss_bwd": [c_p, c_int, c_i64, c_int, c_int, c_f, c_f, c_p, c_p, c_int, c_p],
    "jg_gather_rows": [c_p, c_int, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_gather_rows_bwd": [c_p, c_int, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "jg_l2norm_fwd": [c_p, c_int, c_p, c_p, c_i64, c_int, c_f, c_p],
    "jg_l2norm_bwd": [c_p, c_p, c_p, c_p, c_int, c_i64, c_int, c_f, c_p],
    "jg_patch_nce_fwd": [c_p, c_p, c_int, c_int, c_int, c_f, c_p, c_p, c_p],
    "jg_patch_nce_bwd": [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_p, c_p, c_p],
    "jg_monce_fwd": [c_p, c_p, c_int, c_int, c_int, c_f, c_int, c_int, c_p, c_p, c_p, c_p],
    "jg_monce_bwd": [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_int, c_int, c_p, c_p, c_p, c_p],
}
_U64_FUNCS = {"jg_kernel_launches": []}
_SIZE_T_FUNCS = {
    "jg_groupnorm_fwd_ws_floats": [c_int, c_int, c_int],
    "jg_groupnorm_bwd_ws_floats": [c_int, c_int, c_int],
    "jg_monce_ws_floats": [c_int, c_int, c_int, c_int],
}


def exported_symbols():
    """Every symbol include/jg_b200.h declares (used by the CPU test that checks the .so exports)."""
    return sorted(list(_SIGNATURES.keys()) + list(_SIZE_T_FUNCS.keys()) + list(_U64_FUNCS.keys()) + ["jg_last_error"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libjg_b200.so not found at %s — run `python -m joligen_b200.build` (there is no fallback path)"
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.jg_last_error.restype = ctypes.c_char_p
    lib.jg_last_error.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    for name, argtypes in _SIZE_T_FUNCS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_size_t
        fn.argtypes = argtypes
    for name, argtypes in _U64_FUNCS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_ulonglong
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().jg_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


# kernels launched through the library since import (exact: see jg_kernel_launches in include/jg_b200.h)
launch_count = [0]
call_hook = [None]  # optional profiling hook: fn(name, args) -> context manager


def call(name, *args):
    lib = load()
    before = lib.jg_kernel_launches()
    hook = call_hook[0]
    if hook is None:
        _check(getattr(lib, name)(*args), name)
    else:
        with hook(name, args):
            _check(getattr(lib, name)(*args), name)
    launch_count[0] += lib.jg_kernel_launches() - before  # exact: every launch site of the library counts itself
