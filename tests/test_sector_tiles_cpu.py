"""CPU: the design fact the default tensor-core path rests on — grouping reference pixels by the angle of
(pixel - epipole) makes a 32-pixel tile touch ~2.4x fewer source pixels than a 4x8 block.
numpy restatement of sector_order_kernel's key (csrc/epi_fusion_tile.cu) on the oracle's sample locations."""
import numpy as np

from epipolar_transformers_b200 import config, synthetic as syn
from oracle import epipolar_oracle as eo


def _tap_origin(locs, H, W):
    ix = ((locs[..., 0] + 1) * W - 1) / 2
    iy = ((locs[..., 1] + 1) * H - 1) / 2
    return np.floor(ix).astype(int), np.floor(iy).astype(int)


def _union(x0, y0, sel, H, W):
    m = np.zeros(H * W, bool)
    for dx in (0, 1):
        for dy in (0, 1):
            xx = (x0[:, sel] + dx).ravel(); yy = (y0[:, sel] + dy).ravel()
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            m[yy[ok] * W + xx[ok]] = True
    return int(m.sum())


def sector_order(P_ref, P_src, xs, ys):
    """pixels sorted by (16-bit angle key around e1 = P_ref·C_src, pixel index) — as the device kernel does"""
    C2 = -np.linalg.solve(P_src[:, :3], P_src[:, 3])
    e = P_ref @ np.append(C2, 1.0)
    ex, ey = e[0] / e[2], e[1] / e[2]
    gx, gy = np.meshgrid(xs, ys)
    a0 = np.arctan2(0.5 * (ys[0] + ys[-1]) - ey, 0.5 * (xs[0] + xs[-1]) - ex)
    ang = np.arctan2(gy - ey, gx - ex).ravel() - a0
    ang = (ang + np.pi) % (2 * np.pi) - np.pi
    key = np.minimum(((ang + np.pi) / (2 * np.pi) * 65536).astype(np.int64), 65535)
    return np.argsort(key * 16384 + np.arange(key.size), kind="stable")


def test_sector_tiles_shrink_the_union():
    cfg = config.cfg_h36m_r50_256()
    H = W = 64; K = 64; N = 4
    P1, P2 = syn.pairs_from_ring(N, 4 * H)
    locs = eo.sample_locs(cfg, P1.astype(np.float32), P2.astype(np.float32), H, W, K, dtype=np.float64, geometry="hinf")
    x0, y0 = _tap_origin(locs, H, W)
    xs, ys = eo.pixel_axes(cfg, H, W)
    d_sector, d_block = [], []
    for n in range(N):
        xn, yn = x0[:, n].reshape(K, -1), y0[:, n].reshape(K, -1)
        order = sector_order(P1[n], P2[n], xs, ys)
        assert sorted(order.tolist()) == list(range(H * W))                 # a permutation: every pixel in exactly one tile
        d_sector += [_union(xn, yn, order[i:i + 32], H, W) for i in range(0, H * W, 32)]
        for ty in range(0, H, 4):
            for tx in range(0, W, 8):
                sel = (np.arange(ty, ty + 4)[:, None] * W + np.arange(tx, tx + 8)[None]).ravel()
                d_block.append(_union(xn, yn, sel, H, W))
    d_sector, d_block = np.array(d_sector), np.array(d_block)
    assert d_sector.max() <= 480, d_sector.max()            # the kernel's DMAX: sector tiles never need splitting here
    assert d_sector.mean() < 0.5 * d_block.mean(), (d_sector.mean(), d_block.mean())
    assert d_sector.mean() < 200
