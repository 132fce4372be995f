"""TEST / BASELINE INFRASTRUCTURE — CPU port of the reference's op composition.  NOT product code.

The reference is pure PyTorch and cannot travel to the GPU box, so the CPU
baseline that bench.py reports ("cpu_baseline", `--impl reference`) is this port: it issues the
SAME ATen operators in the same order and with the same tensor shapes as
/root/reference/modeling/layers/epipolar.py:188-255 — per batch item two F.grid_sample calls on
the K-expanded source view (:199,:210), mul+sum over C (:295), ==0 mask (:298), scale+softmax
(:306-307), argmax+gather (:237-241), mul+sum over K (:243), then conv1x1 + BN + residual
(:249-253) — so its run time is the reference's run time on the same host cores.  It is also
checked against the golden vectors (tests/test_oracle_golden.py), so it doubles as a third oracle.
Sample locations come from oracle/epipolar_oracle.py ('reference' pinv geometry, <1 % of the time).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import epipolar_oracle as eo


def forward(cfg, feat_ref, feat_src, P_ref, P_src, params=None, locs=None, threads=None):
    """feat_*: torch float32 [N,C,H,W] (CPU for the baseline; a CUDA tensor runs the same ATen op sequence on the
    GPU, which is the "reference PyTorch forward on the same B200" of BASELINE.md B2); P_*: numpy [N,3,4].
    Returns (finalout, corr_pos, attn)."""
    if threads:
        torch.set_num_threads(int(threads))
    N, C, H, W = feat_ref.shape
    K = int(cfg.EPIPOLAR.SAMPLESIZE)
    scale = float(cfg.EPIPOLAR.SOFTMAXSCALE)
    correct = bool(cfg.EPIPOLAR.USE_CORRECT_NORMALIZE)
    with torch.no_grad():
        if locs is None:
            locs = eo.sample_locs(cfg, np.asarray(P_ref, np.float32), np.asarray(P_src, np.float32), H, W, K,
                                  dtype=np.float32, geometry="reference")
        grid = torch.as_tensor(np.asarray(locs), dtype=torch.float32).to(feat_ref.device)   # [K,N,H,W,2]
        src_k = feat_src.unsqueeze(0).expand(K, N, C, H, W)                      # stride-0 view over K
        fused, corr, weights = [], [], []
        for n in range(N):
            g = grid[:, n]
            keys = F.grid_sample(src_k[:, n], g, align_corners=False)            # [K,C,H,W]
            vals = F.grid_sample(src_k[:, n], g, align_corners=False)            # the reference samples twice (fact 3)
            sim = (keys * feat_ref[n].unsqueeze(0).expand(K, -1, -1, -1)).sum(1)
            sim[sim == 0] = -1e10
            sim = F.softmax(sim * scale, 0)
            top = sim.argmax(0)
            pos = torch.gather(g, 0, top.view(1, H, W, 1).expand(-1, -1, -1, 2)).squeeze(0)
            wh = torch.tensor([W, H], dtype=pos.dtype, device=pos.device)
            corr.append((pos + 1) * (wh - 1) / 2.0 if correct else (pos + 1) * wh / 2.0 - 0.5)      # de_normalize, multiview.py:39-57
            fused.append((vals * sim.view(K, 1, H, W)).sum(0))
            weights.append(sim)
        out = torch.stack(fused)
        if "z" in cfg.EPIPOLAR.PARAMETERIZED:
            p = {k: torch.as_tensor(v).to(out.device) for k, v in params.items()}
            y = F.conv2d(out, p["z.weight"], p["z.bias"])
            y = F.batch_norm(y, p["bn.running_mean"], p["bn.running_var"], p["bn.weight"], p["bn.bias"], False, 0.1, 1e-5)
            out = y + out if cfg.EPIPOLAR.ZRESIDUAL else y
        return out, torch.stack(corr), torch.stack(weights)
