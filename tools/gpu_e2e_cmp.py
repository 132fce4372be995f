"""Developer tool: e2e through HostStreamer with and without the dedicated D2H stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epipolar_transformers_b200 as epi
from epipolar_transformers_b200 import synthetic as syn
dev = torch.device("cuda", 0)
N, C, H, W, K = 4, 256, 64, 64, 64
cfg = epi.cfg_h36m_r50_256()
m = epi.Epipolar(cfg=cfg).to(dev).eval()
m.load_state_dict({k: torch.from_numpy(v) for k, v in syn.z_bn_params(C).items()}, strict=False)
P1, P2 = syn.pairs_from_ring(N, 4 * H)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(syn.features(N, C, H, W, "relu_smooth", s)) for s in (1, 2)]
hp = [pin(P1.astype(np.float32)), pin(P2.astype(np.float32))]
ho = (torch.empty(N, C, H, W).pin_memory(), torch.empty(N, K, H, W).pin_memory(), torch.empty(N, H, W, 2).pin_memory())
for d2h in (False, True):
    hs = epi.HostStreamer(m, dev, depth=2, d2h_stream=d2h)
    for _ in range(5): hs(h[0], h[1], hp[0], hp[1], *ho)
    hs.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): hs(h[0], h[1], hp[0], hp[1], *ho)
    hs.synchronize()
    print("d2h_stream=%s: %.3f ms/step" % (d2h, (time.perf_counter() - t0) * 1e3 / 50))
