"""find_tensor_peak_batch on the GPU — the step right after the fusion layer's 1x1 head.

Mirrors /root/reference/modeling/backbones/basic_batch.py:17-63 (same name, arguments and return value for a
[J,H,W] heat-map) and adds the batched form the caller's Python loop (modeling/backbones/resnet.py:423-428) needs:
a [B,J,H,W] stack in one launch.  All arithmetic runs in libepipolar_b200.so (csrc/epi_peaks.cu); no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def find_tensor_peak_batch(heatmap: torch.Tensor, radius, downsample, threshold: float = 0.000001, int_div: bool = False):
    """heatmap [J,H,W] -> (locs [J,2] (x, y), score [J]);  heatmap [B,J,H,W] -> ([B,J,2], [B,J]).

    int_div=False reproduces what the reference computes under current torch (`index / W` is a true division,
    basic_batch.py:26); int_div=True is the integer division of the torch < 1.4 the reference's README targets."""
    lib = _lib.load()
    if not isinstance(heatmap, torch.Tensor) or heatmap.dim() not in (3, 4):
        raise ValueError("The dimension of the heatmap is wrong : %s" % (tuple(heatmap.shape),))
    if not heatmap.is_cuda:
        raise RuntimeError("heatmap is on %s: the B200 peak finder has no CPU implementation" % heatmap.device)
    if not (radius > 0):
        raise ValueError("The radius is not ok : %r" % (radius,))
    batched = heatmap.dim() == 4
    h = heatmap if batched else heatmap.unsqueeze(0)
    h = h.detach().to(torch.float32).contiguous()
    B, J, H, W = h.shape
    if H <= 1 or W <= 1:
        raise ValueError("To avoid the normalization function divide zero")
    locs = torch.empty((B, J, 2), device=h.device, dtype=torch.float32)
    score = torch.empty((B, J), device=h.device, dtype=torch.float32)
    with torch.cuda.device(h.device):
        stream = torch.cuda.current_stream(h.device).cuda_stream
        _lib.check(lib.epi_find_peaks_f32(h.data_ptr(), locs.data_ptr(), score.data_ptr(), B, J, H, W, float(radius),
                                          float(downsample), float(threshold), int(bool(int_div)), ctypes.c_void_p(stream)),
                   "epi_find_peaks_f32")
    return (locs, score) if batched else (locs[0], score[0])
