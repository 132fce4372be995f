"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/epi_oracle.c (scalar C restatement)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libepi_oracle.so")
_SRC = os.path.join(_HERE, "epi_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.epi_oracle_forward.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def forward(cfg, feat_ref, feat_src, P_ref, P_src, locs=None, align_corners=False, geom_fp32=False,
            threads=0, want_locs=True):
    """Pre-z fusion output of the graded path.  P_* are taken at the precision given
    (pass float32-rounded values to mirror modeling/model.py:183-195).  Returns dict."""
    N, C, H, W = feat_ref.shape
    K = int(cfg.EPIPOLAR.SAMPLESIZE)
    fr = np.ascontiguousarray(feat_ref, np.float32)
    fs = np.ascontiguousarray(feat_src, np.float32)
    P1 = np.ascontiguousarray(P_ref, np.float64)
    P2 = np.ascontiguousarray(P_src, np.float64)
    li = np.ascontiguousarray(locs, np.float32) if locs is not None else None
    out = np.empty((N, C, H, W), np.float32)
    attn = np.empty((N, K, H, W), np.float32)
    corr = np.empty((N, H, W, 2), np.float32)
    lo = np.empty((K, N, H, W, 2), np.float32) if want_locs else None
    f, d = ctypes.c_float, ctypes.c_double
    rc = lib().epi_oracle_forward(
        _p(fr, f), _p(fs, f), _p(P1, d), _p(P2, d), _p(li, f),
        ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(K),
        d(float(cfg.BACKBONE.DOWNSAMPLE)), d(float(cfg.DATASETS.IMAGE_RESIZE * cfg.DATASETS.PREDICT_RESIZE)),
        d(float(cfg.EPIPOLAR.SOFTMAXSCALE)), ctypes.c_int(int(bool(cfg.EPIPOLAR.USE_CORRECT_NORMALIZE))),
        ctypes.c_int(int(bool(align_corners))), ctypes.c_int(int(bool(geom_fp32))),
        _p(out, f), _p(attn, f), _p(corr, f), _p(lo, f), ctypes.c_int(int(threads)))
    if rc != 0:
        raise RuntimeError("epi_oracle_forward rc=%d" % rc)
    return {"out": out, "attn": attn, "corr_pos": corr, "sample_locs": lo}
