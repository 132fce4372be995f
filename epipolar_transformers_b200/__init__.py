"""B200-native epipolar-transformer fusion path (drop-in for the reference's
modeling/layers/epipolar.py::Epipolar + the projection helpers of vision/multiview.py).

    from epipolar_transformers_b200 import Epipolar, set_global_cfg
    sampler = Epipolar()                       # reads the global cfg like the reference
    out, corr_pos, attn, locs = sampler(feat_ref, feat_src, KRT_ref, KRT_src)

The arithmetic runs in libepipolar_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/epipolar_b200.h).  Importing this package does not need a GPU; calling the op does,
and fails loudly if the library is missing.
"""
from .config import Node, default_cfg, make_cfg, get_global_cfg, set_global_cfg, cfg_h36m_r50_256, cfg_h36m_r152_384
from .epipolar import Epipolar, FusionState, ZeroInitBN, epipolar_fusion, fold_z_bn, sample_locs, fused_other_feat
from .host_pipeline import HostStreamer, bind_host_to_gpu
from .peaks import find_tensor_peak_batch
from . import multiview, synthetic

__all__ = ["Epipolar", "FusionState", "HostStreamer", "bind_host_to_gpu", "ZeroInitBN", "epipolar_fusion", "fold_z_bn", "sample_locs", "fused_other_feat", "find_tensor_peak_batch",
           "Node", "default_cfg", "make_cfg", "get_global_cfg", "set_global_cfg",
           "cfg_h36m_r50_256", "cfg_h36m_r152_384", "multiview", "synthetic"]
