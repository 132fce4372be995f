"""Developer tool: fused-forward time of the pipelined tensor-core kernel vs the CUDA-core warp kernel over map sizes (auto-selection threshold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epipolar_transformers_b200 as epi
from epipolar_transformers_b200 import synthetic as syn
C = 256
for (H, K) in ((64, 64), (96, 64), (112, 64), (128, 64), (160, 64), (256, 64), (128, 48), (256, 48), (128, 32), (256, 32), (256, 16), (64, 128)):
    W, N = H, 2
    P1, P2 = syn.pairs_from_ring(N, 4 * H)
    P1 = torch.from_numpy(P1.astype(np.float32)).cuda(); P2 = torch.from_numpy(P2.astype(np.float32)).cuda()
    f1 = torch.relu(torch.randn(N, C, H, W, device="cuda")); f2 = torch.relu(torch.randn(N, C, H, W, device="cuda"))
    row = []
    for var in ("pipe", "warp", "auto"):
        st = epi.FusionState()
        try:
            for _ in range(3): epi.epipolar_fusion(f1, f2, P1, P2, K=K, correct_normalize=True, variant=var, state=st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): epi.epipolar_fusion(f1, f2, P1, P2, K=K, correct_normalize=True, variant=var, state=st)
            e1.record(); torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) / 5)
        except Exception as ex:
            row.append(float("nan"))
    print("H=W=%d K=%d C=%d N=%d: pipe %.3f ms, warp %.3f ms, auto %.3f ms  (ns/px: %.1f / %.1f / %.1f)" % (H, K, C, N, row[0], row[1], row[2], row[0] * 1e6 / (N * H * W), row[1] * 1e6 / (N * H * W), row[2] * 1e6 / (N * H * W)))
