"""B200-native CenterFace inference path (sm_100a CUDA kernels behind a C ABI).

The package directory name contains hyphens (it mirrors the reference repository's name), so it
is imported with ``importlib.import_module("lightweight-face-detection-centernet_b200")``; the
repo-root ``__graft_entry__.py`` shows how.
"""
from . import _lib
from ._lib import (CF_DECODE_A, CF_DECODE_B, CF_IN_F32_NCHW, CF_IN_U8_HWC, CF_PW_SIMT, CF_PW_TCGEN05,
                   CF_PW_TCGEN05_1P, CF_PW_TCGEN05_LAYERWISE, CF_PW_TCGEN05_MIXED, CenterFaceError)
from .build import build
from .weights import load_state_dict, pack_weights

__all__ = ["CenterFace", "CenterFaceNet", "Engine", "ctdet_decode", "decode_threshold", "get_detections", "ctdet_post_process", "resize_u8", "warp_affine_u8", "letterbox_u8", "letterbox_matrix", "evaluate", "bbox_overlap",
           "write_detections_txt", "build",
           "pack_weights", "load_state_dict", "CenterFaceError", "CF_DECODE_A", "CF_DECODE_B", "CF_IN_F32_NCHW",
           "CF_IN_U8_HWC", "CF_PW_SIMT", "CF_PW_TCGEN05", "CF_PW_TCGEN05_1P", "CF_PW_TCGEN05_LAYERWISE", "CF_PW_TCGEN05_MIXED"]


def __getattr__(name):  # engine/centerface import torch lazily; keep `import pkg` light
    if name in ("CenterFace", "CenterFaceNet", "get_detections"):
        from . import centerface as m
        return getattr(m, name)
    if name in ("evaluate", "bbox_overlap", "write_detections_txt"):
        from . import widerface as m
        return getattr(m, name)
    if name in ("Engine", "ctdet_decode", "decode_threshold", "ctdet_post_process", "inverse_affine", "resize_u8", "warp_affine_u8", "letterbox_u8",
                "letterbox_matrix"):
        from . import engine as m
        return getattr(m, name)
    raise AttributeError(name)
