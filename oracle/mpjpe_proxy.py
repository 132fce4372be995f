"""TEST INFRASTRUCTURE — synthetic MPJPE proxy.  NOT product code.

H36M and the released weights are not available offline, so the north star's accuracy criterion
("MPJPE within 0.1 mm of the reference on identical inputs") is evaluated on a synthetic stand-in of the
reference's test path:  3-D joints -> per-view Gaussian heat-map features (+ distractor channels) ->
epipolar fusion of every view with its nearest view (the layer under test) + the caller residual
(/root/reference/modeling/backbones/resnet.py:388) -> fixed 1x1 head (resnet.py:421) -> sub-pixel peak
(soft-argmax in a window around the arg-max, the role of find_tensor_peak_batch,
modeling/backbones/basic_batch.py:17-63) -> linear (DLT) triangulation over the views
(vision/multi_camera_system.py:199-225 does the same SVD) -> mean per-joint position error
(modeling/metrics/metrics3d.py:5-46).  Everything downstream of the fusion is the same numpy code for the
reference run and for the CUDA run, so the difference in MPJPE is attributable to the fusion layer.
"""
from __future__ import annotations

import numpy as np

from epipolar_transformers_b200 import config, synthetic as syn

V, J, C, H, W, K, IMG = 4, 17, 64, 64, 64, 64, 256
SIGMA = 1.6          # heat-map sigma in feature px


def build(seed=0):
    rng = np.random.default_rng(seed)
    KRT = syn.ring_cameras(V, IMG)
    joints = np.array([0.0, 0.0, 1000.0]) + rng.uniform(-450, 450, size=(J, 3))        # mm, around the look-at point
    uv = np.stack([(P @ np.concatenate([joints, np.ones((J, 1))], 1).T).T for P in KRT])   # [V,J,3]
    uv = uv[..., :2] / uv[..., 2:3]                                                      # image px
    fpx = (uv + 0.5 - 2.0) / 4.0                                                         # coord2pix, DOWNSAMPLE=4
    ys, xs = np.mgrid[0:H, 0:W]
    feats = np.zeros((V, C, H, W), np.float32)
    for v in range(V):
        for j in range(J):
            feats[v, j] = np.exp(-((xs - fpx[v, j, 0]) ** 2 + (ys - fpx[v, j, 1]) ** 2) / (2 * SIGMA ** 2))
        feats[v, :J] += 0.15 * np.maximum(rng.standard_normal((J, H, W)), 0).astype(np.float32)   # detector noise
        feats[v, J:] = np.maximum(rng.standard_normal((C - J, H, W)), 0).astype(np.float32) * 0.3  # distractor channels
    src = syn.nearest_source(KRT)
    cfg = config.make_cfg(KEYPOINT=dict(HEATMAP_SIZE=(H, W), NFEATS=C),
                          EPIPOLAR=dict(SAMPLESIZE=K, USE_CORRECT_NORMALIZE=True))
    head = np.zeros((J, C), np.float32)
    head[np.arange(J), np.arange(J)] = 1.0
    head += 0.02 * rng.standard_normal((J, C)).astype(np.float32)
    return dict(cfg=cfg, KRT=KRT, joints=joints, feat_ref=feats, feat_src=feats[src].copy(),
                P_ref=KRT.astype(np.float32), P_src=KRT[src].astype(np.float32), head=head)


def _peaks(heat):
    """sub-pixel peak per joint: arg-max then intensity centroid in a 7x7 window -> feature px (x, y)."""
    Jn, Hh, Ww = heat.shape
    out = np.zeros((Jn, 2))
    for j in range(Jn):
        y0, x0 = np.unravel_index(np.argmax(heat[j]), (Hh, Ww))
        ya, yb, xa, xb = max(0, y0 - 3), min(Hh, y0 + 4), max(0, x0 - 3), min(Ww, x0 + 4)
        win = np.maximum(heat[j, ya:yb, xa:xb].astype(np.float64), 0) ** 2
        yy, xx = np.mgrid[ya:yb, xa:xb]
        out[j] = [(win * xx).sum() / win.sum(), (win * yy).sum() / win.sum()]
    return out


def mpjpe(d, fused):
    """fused: [V,C,H,W] output of the fusion layer (finalout).  Returns MPJPE in mm."""
    x = fused.astype(np.float64) + d["feat_ref"].astype(np.float64)            # caller residual ret + feat
    pts2d = []
    for v in range(V):
        heat = np.einsum("jc,chw->jhw", d["head"].astype(np.float64), x[v])
        pts2d.append(_peaks(heat) * 4.0 + 2.0 - 0.5)                            # pix2coord -> image px
    pts2d = np.stack(pts2d)                                                     # [V,J,2]
    errs = []
    for j in range(J):
        A = []
        for v in range(V):
            P = d["KRT"][v]
            A.append(pts2d[v, j, 0] * P[2] - P[0]); A.append(pts2d[v, j, 1] * P[2] - P[1])
        _, _, vt = np.linalg.svd(np.stack(A))
        X = vt[-1, :3] / vt[-1, 3]
        errs.append(np.linalg.norm(X - d["joints"][j]))
    return float(np.mean(errs))
