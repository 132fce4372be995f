"""TEST INFRASTRUCTURE — freezes outputs of the UNMODIFIED reference find_tensor_peak_batch
(/root/reference/modeling/backbones/basic_batch.py:17-63) as tests/golden/peaks.npz.  Build container only.
    python -m oracle.make_golden_peaks
Inputs are regenerated from seeds by `cases()` (also used by the tests)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {   # name: (J, H, W, radius, downsample, kind)
    "h36m_r50": (17, 64, 64, 8.0, 4.0, "gauss"),            # KEYPOINT.SIGMA = 8 (keypoint_h36m_zresidual_fixed.yaml:40)
    "r152_384": (17, 96, 96, 8.0, 4.0, "gauss"),
    "small_radius": (5, 20, 28, 2.0, 8.0, "gauss"),
    "border": (6, 16, 16, 3.0, 4.0, "border"),              # peaks on the map border: zero padding of the window
    "noise": (8, 24, 24, 4.0, 4.0, "noise"),
}


def heatmaps(name):
    J, H, W, radius, ds, kind = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.zeros((J, H, W), np.float32)
    for j in range(J):
        if kind == "border":
            cx, cy = [(0.0, 0.0), (W - 1.0, 0.3), (0.4, H - 1.0), (W - 1.0, H - 1.0), (W / 2, 0.0), (0.0, H / 2)][j % 6]
        else:
            cx, cy = rng.uniform(2, W - 3), rng.uniform(2, H - 3)
        g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * (1.0 + 0.25 * radius) ** 2))
        if kind == "noise":
            g = 0.2 * g + rng.random((H, W)) * 0.3
        else:
            g = g + 0.01 * rng.standard_normal((H, W))
        out[j] = g.astype(np.float32)
    return out


def main():
    import torch
    from tests.test_reference_dropin_cpu import _import_reference_resnet
    _import_reference_resnet()
    import importlib
    bb = importlib.import_module("modeling.backbones.basic_batch")
    rec = {}
    for name, (J, H, W, radius, ds, kind) in CASES.items():
        h = heatmaps(name)
        locs, score = bb.find_tensor_peak_batch(torch.from_numpy(h), radius, ds)
        rec[name + "_locs"] = locs.numpy(); rec[name + "_score"] = score.numpy()
        print(name, locs[:2].numpy().round(3).tolist())
    rec["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "peaks.npz"), **rec)


if __name__ == "__main__":
    main()
