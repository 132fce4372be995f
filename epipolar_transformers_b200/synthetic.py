"""Seeded synthetic inputs for the epipolar fusion path.

Everything is generated with numpy's PCG64 `default_rng(seed)` so the same arrays are
produced in the build container (where the golden vectors are made from the reference)
and on the GPU box (where only the seeds travel for the larger shapes).

Cameras mirror how the reference's dataset builds KRT = K [R | -R C] in float64 and casts
to float32 later (/root/reference/data/datasets/joints_dataset.py:334-336,
/root/reference/modeling/model.py:183-195); the source view of a reference view is its
nearest camera centre (/root/reference/vision/multiview.py:59-83).
"""
from __future__ import annotations

import numpy as np


def ring_cameras(n_views: int = 4, img_size: int = 256, radius: float = 5000.0,
                 height: float = 1500.0, target=(0.0, 0.0, 1000.0), jitter: float = 0.0,
                 seed: int = 0) -> np.ndarray:
    """H36M-like ring: returns KRT [V,3,4] float64. f = 290*(img/256), c = img/2, world z up."""
    rng = np.random.default_rng(seed)
    f = 290.0 * (img_size / 256.0)
    c = img_size / 2.0
    Kmat = np.array([[f, 0, c], [0, f, c], [0, 0, 1.0]])
    target = np.asarray(target, dtype=np.float64)
    out = np.zeros((n_views, 3, 4))
    for v in range(n_views):
        ang = 2.0 * np.pi * v / n_views + 0.3
        C = np.array([radius * np.cos(ang), radius * np.sin(ang), height])
        if jitter > 0:
            C = C + rng.normal(size=3) * jitter
        fwd = target - C
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])            # world -> camera (x right, y down, z fwd)
        out[v] = Kmat @ np.concatenate([R, (-R @ C)[:, None]], axis=1)
    return out


def camera_centers(KRT: np.ndarray) -> np.ndarray:
    """C = -A^-1 t for each [3,4] (vision/multiview.py:13-15)."""
    return np.stack([-np.linalg.solve(P[:, :3], P[:, 3]) for P in KRT])


def nearest_source(KRT: np.ndarray) -> np.ndarray:
    """src(v) = nearest other camera centre (vision/multiview.py:59-83, TOPK=1)."""
    C = camera_centers(KRT)
    d = np.linalg.norm(C[:, None] - C[None], axis=-1)
    np.fill_diagonal(d, np.inf)
    return d.argmin(1)


def pairs_from_ring(n_pairs: int, img_size: int, seed: int = 0, jitter: float = 0.0):
    """(P_ref, P_src) float64 [N,3,4]: view v paired with its nearest neighbour."""
    KRT = ring_cameras(n_pairs, img_size, seed=seed, jitter=jitter)
    src = nearest_source(KRT)
    return KRT.copy(), KRT[src].copy()


def random_krt(n_pairs: int, seed: int = 0):
    """Literal randn KRTs (BASELINE config 1: 'random KRT'); legal in the reference."""
    rng = np.random.default_rng(seed + 7919)
    return rng.standard_normal((n_pairs, 3, 4)), rng.standard_normal((n_pairs, 3, 4))


def features(N: int, C: int, H: int, W: int, kind: str = "randn", seed: int = 0) -> np.ndarray:
    """float32 [N,C,H,W].  'randn': worst-case sensitivity.  'relu_smooth': post-ReLU,
    low-frequency maps like a deconv head's output (resnet.py:358-359)."""
    rng = np.random.default_rng(seed)
    if kind == "randn":
        return rng.standard_normal((N, C, H, W), dtype=np.float32)
    if kind == "relu_smooth":
        h8, w8 = max(2, H // 8), max(2, W // 8)
        coarse = rng.standard_normal((N, C, h8, w8))
        ys = np.linspace(0, h8 - 1, H)
        xs = np.linspace(0, w8 - 1, W)
        y0 = np.clip(np.floor(ys).astype(int), 0, h8 - 2)
        x0 = np.clip(np.floor(xs).astype(int), 0, w8 - 2)
        fy = (ys - y0)[None, None, :, None]
        fx = (xs - x0)[None, None, None, :]
        a = coarse[:, :, y0][:, :, :, x0]
        b = coarse[:, :, y0][:, :, :, x0 + 1]
        c = coarse[:, :, y0 + 1][:, :, :, x0]
        d = coarse[:, :, y0 + 1][:, :, :, x0 + 1]
        up = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
        return np.maximum(up, 0.0).astype(np.float32)
    raise ValueError(kind)


def z_bn_params(C: int, seed: int = 0):
    """Non-trivial z conv + BN(eval) parameters (zero-init BN would hide the branch).
    Returns dict of float32 arrays with the reference's parameter names (epipolar.py:64-65)."""
    rng = np.random.default_rng(seed + 104729)
    return {
        "z.weight": (rng.standard_normal((C, C, 1, 1)) / np.sqrt(C)).astype(np.float32),
        "z.bias": (rng.standard_normal(C) * 0.1).astype(np.float32),
        "bn.weight": (rng.standard_normal(C) * 0.5).astype(np.float32),
        "bn.bias": (rng.standard_normal(C) * 0.1).astype(np.float32),
        "bn.running_mean": (rng.standard_normal(C) * 0.2).astype(np.float32),
        "bn.running_var": (rng.uniform(0.5, 1.5, C)).astype(np.float32),
    }
