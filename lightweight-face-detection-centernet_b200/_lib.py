"""ctypes binding of include/centerface_b200.h -- the stub a maintainer of the reference would
add.  There is no fallback: if the shared library is missing or a call fails,
an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcenterface_b200.so")

CF_IN_F32_NCHW, CF_IN_U8_HWC = 0, 1
CF_PW_SIMT, CF_PW_TCGEN05, CF_PW_TCGEN05_1P, CF_PW_TCGEN05_LAYERWISE, CF_PW_TCGEN05_MIXED = 0, 1, 2, 3, 6
CF_DECODE_A, CF_DECODE_B = 0, 1
CLS_ALL, CLS_PW, CLS_DW, CLS_STEM, CLS_HEADS, CLS_DECODE, CLS_FUSED = range(7)
MAX_CAP = 4096

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/centerface_b200.h
SIGNATURES = {
    "cf_create": (C.c_int, [_vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "cf_destroy": (C.c_int, [_vp]),
    "cf_last_error": (C.c_char_p, []),
    "cf_abi_version": (C.c_int, []),
    "cf_weights_blob_bytes": (C.c_size_t, []),
    "cf_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "cf_heads": (C.c_int, [_vp] + [C.POINTER(_vp)] * 5),
    "cf_tap": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), _i, _i, _i]),
    "cf_ctdet_decode": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "cf_ctdet_decode_classes": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "cf_decode_topk": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "cf_ctdet_post_process": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "cf_decode_threshold": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                      C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _vp, _vp, _vp, _vp]),
    "cf_detect_topk_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "cf_submit_topk_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "cf_wait_host": (C.c_int, [_vp]),
    "cf_detect_threshold_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_int, _vp, _vp, _vp]),
    "cf_debug_pw_gemm": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "cf_debug_pw_gemm_time": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int,
                                        C.POINTER(C.c_float), C.c_char_p, C.c_int]),
    "cf_debug_swish": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int]),
    "cf_debug_tma_stream": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "cf_resize_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _i]),
    "cf_resize_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "cf_warp_affine_tables": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_size_t]),
    "cf_warp_affine_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp]),
    "cf_detect_image_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_int, _vp, _vp, _vp]),
    "cf_launch_count": (C.c_longlong, [_vp]),
    "cf_fused_block_mask": (C.c_uint, [C.c_int]),
    "cf_dwp_block_mask": (C.c_uint, [C.c_int]),
    "cf_comm_unique_id": (C.c_int, [_vp]),
    "cf_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "cf_submit_topk_gather_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "cf_nms": (C.c_int, [_vp, _vp, C.c_int, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    "cf_nms_scratch_bytes": (C.c_size_t, [C.c_int]),
    "cf_nms_host": (C.c_int, [C.c_int, _vp, _vp, C.c_int, C.c_float, _vp, _vp]),
    "cf_debug_mbf_trace": (C.c_int, [_vp, _vp, C.c_int]),
    "cf_work_model": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cf_replay_class": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "cf_time_class": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_float), _i]),
    "cf_time_steps": (C.c_int, [_vp, C.c_int, _vp, C.POINTER(C.c_float), _i, C.c_int, _i]),
}

_lib = None


class CenterFaceError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library (built by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CenterFaceError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cf_last_error().decode(errors="replace")
        raise CenterFaceError(f"{what or 'centerface_b200'} failed ({rc}): {msg}")


def fused_blocks(pw_engine=CF_PW_TCGEN05):
    """Indices of the MBConv blocks that run as one fused kernel under `pw_engine` -- cf_fused_block_mask."""
    m = load().cf_fused_block_mask(pw_engine)
    return [i for i in range(12) if (m >> i) & 1]


def dwp_blocks(pw_engine=CF_PW_TCGEN05):
    """Indices of the blocks whose depth-wise + projection run as one kernel under `pw_engine` -- cf_dwp_block_mask."""
    m = load().cf_dwp_block_mask(pw_engine)
    return [i for i in range(12) if (m >> i) & 1]


def work_model(h, w, in_format=CF_IN_U8_HWC, which=CLS_ALL, pw_engine=CF_PW_TCGEN05):
    """(bytes, flops) of one image -- cf_work_model."""
    b, f = C.c_double(), C.c_double()
    check(load().cf_work_model(h, w, in_format, pw_engine, which, C.byref(b), C.byref(f)), "cf_work_model")
    return b.value, f.value
