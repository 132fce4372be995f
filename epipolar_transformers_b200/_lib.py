"""ctypes binding of libepipolar_b200.so (the C ABI in include/epipolar_b200.h).

There is NO fallback: if the CUDA library is missing or does not export the ABI the import of
the op fails loudly (RuntimeError), so a GPU box can never silently run another code path.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libepipolar_b200.so")

EPI_ABI_VERSION = 2
EPI_VARIANT_AUTO, EPI_VARIANT_WARP, EPI_VARIANT_TILE, EPI_VARIANT_SECTOR, EPI_VARIANT_PIPE = 0, 1, 2, 3, 4
VARIANTS = {"auto": EPI_VARIANT_AUTO, "warp": EPI_VARIANT_WARP, "tile": EPI_VARIANT_TILE, "sector": EPI_VARIANT_SECTOR,
            "pipe": EPI_VARIANT_PIPE}

EXPORTS = ("epi_version", "epi_last_error", "epi_fusion_workspace_bytes", "epi_fusion_cache_bytes", "epi_fusion_forward_f32",
           "epi_fusion_backward_workspace_bytes", "epi_fusion_backward_f32", "epi_find_peaks_f32",
           "epi_sample_locs_f32", "epi_fold_z_bn_f32", "epi_last_launch_count", "epi_umma_selftest",
           "epi_kernel_timing_enable", "epi_kernel_timing_last_ms", "epi_kernel_timing_last3")

_fp = ctypes.POINTER(ctypes.c_float)


class EpiFusionParams(ctypes.Structure):
    """Field-for-field mirror of `struct EpiFusionParams` (include/epipolar_b200.h)."""
    _fields_ = [
        ("feat_ref", ctypes.c_void_p), ("ref_stride", ctypes.c_int64 * 4),
        ("feat_src", ctypes.c_void_p), ("src_stride", ctypes.c_int64 * 4),
        ("P_ref", ctypes.c_void_p), ("P_src", ctypes.c_void_p), ("sample_locs_in", ctypes.c_void_p),
        ("out", ctypes.c_void_p), ("out_stride", ctypes.c_int64 * 4),
        ("attn", ctypes.c_void_p), ("corr_pos", ctypes.c_void_p), ("sample_locs_out", ctypes.c_void_p),
        ("z_weight_folded", ctypes.c_void_p), ("z_bias_folded", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("N", ctypes.c_int32), ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("K", ctypes.c_int32),
        ("downsample", ctypes.c_float), ("img_scale", ctypes.c_float), ("eps", ctypes.c_float), ("softmax_scale", ctypes.c_float),
        ("align_corners", ctypes.c_int32), ("correct_normalize", ctypes.c_int32), ("z_residual", ctypes.c_int32),
        ("add_ref_residual", ctypes.c_int32), ("variant", ctypes.c_int32), ("reserved", ctypes.c_int32 * 3),
        ("cache", ctypes.c_void_p), ("cache_bytes", ctypes.c_size_t),
    ]


class EpiFusionBwdParams(ctypes.Structure):
    """Field-for-field mirror of `struct EpiFusionBwdParams` (include/epipolar_b200.h)."""
    _fields_ = [
        ("feat_ref", ctypes.c_void_p), ("ref_stride", ctypes.c_int64 * 4),
        ("feat_src", ctypes.c_void_p), ("src_stride", ctypes.c_int64 * 4),
        ("P_ref", ctypes.c_void_p), ("P_src", ctypes.c_void_p), ("sample_locs_in", ctypes.c_void_p),
        ("attn", ctypes.c_void_p),
        ("grad_out", ctypes.c_void_p), ("gout_stride", ctypes.c_int64 * 4),
        ("grad_attn", ctypes.c_void_p),
        ("grad_ref", ctypes.c_void_p), ("gref_stride", ctypes.c_int64 * 4),
        ("grad_src", ctypes.c_void_p), ("gsrc_stride", ctypes.c_int64 * 4),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("N", ctypes.c_int32), ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("K", ctypes.c_int32),
        ("downsample", ctypes.c_float), ("img_scale", ctypes.c_float), ("eps", ctypes.c_float), ("softmax_scale", ctypes.c_float),
        ("align_corners", ctypes.c_int32), ("correct_normalize", ctypes.c_int32),
        ("grad_keys", ctypes.c_int32), ("grad_vals", ctypes.c_int32), ("reserved", ctypes.c_int32 * 4),
    ]


_lib = None


def load():
    """Load the shared library once; raise if it is absent or the ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "epipolar_transformers_b200: CUDA library %s is missing. Build it with "
            "`python -m epipolar_transformers_b200.build` (needs nvcc). There is no CPU/PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise RuntimeError("libepipolar_b200.so does not export %s" % missing)
    lib.epi_version.restype = ctypes.c_int
    lib.epi_last_error.restype = ctypes.c_char_p
    lib.epi_last_launch_count.restype = ctypes.c_int
    lib.epi_fusion_workspace_bytes.restype = ctypes.c_size_t
    lib.epi_fusion_workspace_bytes.argtypes = [ctypes.POINTER(EpiFusionParams)]
    lib.epi_fusion_cache_bytes.restype = ctypes.c_size_t
    lib.epi_fusion_cache_bytes.argtypes = [ctypes.POINTER(EpiFusionParams)]
    lib.epi_fusion_forward_f32.restype = ctypes.c_int
    lib.epi_fusion_forward_f32.argtypes = [ctypes.POINTER(EpiFusionParams), ctypes.c_void_p]
    lib.epi_fusion_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.epi_fusion_backward_workspace_bytes.argtypes = [ctypes.POINTER(EpiFusionBwdParams)]
    lib.epi_fusion_backward_f32.restype = ctypes.c_int
    lib.epi_fusion_backward_f32.argtypes = [ctypes.POINTER(EpiFusionBwdParams), ctypes.c_void_p]
    lib.epi_find_peaks_f32.restype = ctypes.c_int
    lib.epi_find_peaks_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_void_p]
    lib.epi_sample_locs_f32.restype = ctypes.c_int
    lib.epi_sample_locs_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_int32, ctypes.c_void_p]
    lib.epi_fold_z_bn_f32.restype = ctypes.c_int
    lib.epi_fold_z_bn_f32.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_float, ctypes.c_int32, ctypes.c_void_p,
                                                             ctypes.c_void_p, ctypes.c_void_p]
    lib.epi_umma_selftest.restype = ctypes.c_int
    lib.epi_umma_selftest.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p]
    lib.epi_kernel_timing_enable.restype = ctypes.c_int
    lib.epi_kernel_timing_enable.argtypes = [ctypes.c_int]
    lib.epi_kernel_timing_last_ms.restype = ctypes.c_float
    lib.epi_kernel_timing_last3.restype = ctypes.c_int
    lib.epi_kernel_timing_last3.argtypes = [ctypes.POINTER(ctypes.c_float)]
    v = lib.epi_version()
    if v != EPI_ABI_VERSION:
        raise RuntimeError("libepipolar_b200.so ABI version %d != expected %d" % (v, EPI_ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().epi_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))
