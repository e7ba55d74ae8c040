"""ctypes binding of libvpuformer_b200.so (include/vpuformer_b200.h).

The library is the product: if it is missing or fails to load, importing the compute path
raises -- there is no eager/PyTorch/CPU fallback anywhere in this package.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPU_LIB_PATH") or os.path.join(HERE, "libvpuformer_b200.so")   # env: A/B builds during development

VPU_F32, VPU_BF16, VPU_I32, VPU_F64, VPU_U8 = 0, 1, 2, 3, 4


class VpuDims(ctypes.Structure):
    _fields_ = [("img_size", c_int32), ("patch", c_int32), ("embed_dim", c_int32), ("depth", c_int32),
                ("num_heads", c_int32), ("num_max_points", c_int32), ("dma_depth", c_int32), ("dma_heads", c_int32),
                ("dma_mlp_dim", c_int32), ("ppue_ffn_dim", c_int32), ("head_channels", c_int32),
                ("out_dims", c_int32 * 4), ("norm_radius", c_float)]


class VpuProfileEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("ms", c_double), ("flops", c_double), ("bytes", c_double),
                ("launches", c_int64)]


class VpuPrompts(ctypes.Structure):
    _fields_ = [("points", c_void_p), ("ppue_points", c_void_p), ("n", c_int32), ("n_ppue", c_int32),
                ("type", c_int32), ("boxes", c_void_p), ("scrib_sel", c_void_p), ("scrib_slot", c_void_p),
                ("extra_mask", c_void_p)]


class VpuSessionState(ctypes.Structure):
    _fields_ = [("S", c_int32), ("H", c_int32), ("W", c_int32), ("T", c_int32), ("max_clicks", c_int32), ("n_half", c_int32),
                ("images", c_void_p), ("prev_probs", c_void_p), ("pred", c_void_p), ("clicks", c_void_p), ("nclicks", c_void_p),
                ("roi", c_void_p), ("fgbox", c_void_p), ("pred_thr", c_float), ("zoom_thr", c_float),
                ("expansion_ratio", c_double), ("recompute_thresh_iou", c_double), ("min_crop_size", c_int32)]


# name -> (restype, argtypes); every symbol include/vpuformer_b200.h declares
SIGNATURES = {
    "vpu_last_error": (c_char_p, []),
    "vpu_version": (c_int, []),
    "vpu_launch_count": (ctypes.c_ulonglong, []),
    "vpu_profile_begin": (c_int, [c_void_p]),
    "vpu_profile_end": (c_int, [c_void_p, POINTER(VpuProfileEntry), c_int, POINTER(c_int)]),
    "vpu_create": (c_int, [POINTER(c_void_p), POINTER(VpuDims)]),
    "vpu_destroy": (None, [c_void_p]),
    "vpu_bind_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64), c_int]),
    "vpu_set_scalar": (c_int, [c_void_p, c_char_p, c_float]),
    "vpu_set_click_table": (c_int, [c_void_p, POINTER(c_float), c_int]),
    "vpu_finalize": (c_int, [c_void_p]),
    "vpu_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "vpu_workspace_lookup": (c_int, [c_void_p, c_int, c_char_p, POINTER(c_size_t), POINTER(c_size_t)]),
    "vpu_forward": (c_int, [c_void_p, c_void_p, POINTER(VpuPrompts), c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vpu_ppue": (c_int, [c_void_p, POINTER(VpuPrompts), c_int, c_void_p, c_void_p]),
    "vpu_coord_features": (c_int, [c_void_p, c_void_p, POINTER(VpuPrompts), c_int, c_void_p, c_void_p]),
    "vpu_gemm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                         c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vpu_head_tail": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                              c_void_p, c_void_p]),
    "vpu_gemm_b2b": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vpu_gemm_table": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "vpu_gemm_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                   c_void_p, c_int, c_void_p, c_void_p]),
    "vpu_gemm_pixel_shuffle": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "vpu_attention": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "vpu_noc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vpu_noc_next_clicks": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vpu_raster_prompts": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vpu_session_prepare": (c_int, [POINTER(VpuSessionState), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpu_session_finish": (c_int, [POINTER(VpuSessionState), c_void_p, c_int, c_void_p, c_void_p]),
    "vpu_image_from_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vpu_debug_attention_trace": (c_int, [c_void_p, c_int]),
    "vpu_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
    "vpu_groupnorm_nhwc": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vpu_upsample_align_corners": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_void_p]),
}

_lib = None


class VpuError(RuntimeError):
    pass


def load():
    """Load the CUDA extension; raise loudly if it is not built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VpuError("%s is missing: build it with `python -m pvpuformer_b200.build` "
                       "(there is no CPU/eager fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VpuError("vpuformer_b200 error %d: %s" % (rc, load().vpu_last_error().decode(errors="replace")))


def ptr(t):
    """Device/host pointer of a tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
