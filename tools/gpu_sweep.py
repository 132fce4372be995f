"""Developer tool (GPU box): (1) the reference's ATen op sequence on the same B200 (the >=10x target's
denominator) vs this library, (2) BASELINE config 5 — K x C sweep of the fused path with achieved GB/s."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epipolar_transformers_b200 as epi
from epipolar_transformers_b200 import synthetic as syn
from oracle import torch_port

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
PEAK = 6570.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def ev_time(fn, warm=5, iters=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {"peak_gbs": PEAK, "reference_ops_on_gpu": {}, "sweep": []}
for name, (N, C, H, K, cfgf) in {"cfg2": (4, 256, 64, 64, epi.cfg_h36m_r50_256), "cfg3": (4, 256, 96, 64, epi.cfg_h36m_r152_384)}.items():
    cfg = cfgf(); W = H
    P1, P2 = syn.pairs_from_ring(N, 4 * H)
    f1 = torch.relu(torch.randn(N, C, H, W, device=dev)); f2 = torch.relu(torch.randn(N, C, H, W, device=dev))
    params = syn.z_bn_params(C) if "z" in cfg.EPIPOLAR.PARAMETERIZED else None
    from oracle import epipolar_oracle as eo
    locs = eo.sample_locs(cfg, P1.astype(np.float32), P2.astype(np.float32), H, W, K, dtype=np.float32, geometry="reference")
    t_ref = ev_time(lambda: torch_port.forward(cfg, f1, f2, P1, P2, params=params, locs=locs), warm=3, iters=10)
    m = epi.Epipolar(cfg=cfg).to(dev).eval()
    if params: m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    tP1, tP2 = torch.from_numpy(P1.astype(np.float32)).to(dev), torch.from_numpy(P2.astype(np.float32)).to(dev)
    with torch.no_grad():
        t_ours = ev_time(lambda: m(f1, f2, tP1, tP2), warm=5, iters=50)
    out["reference_ops_on_gpu"][name] = {"reference_ms": t_ref, "ours_ms": t_ours, "speedup": t_ref / t_ours,
                                         "note": "reference = oracle/torch_port.py on CUDA tensors (same ATen op sequence as modeling/layers/epipolar.py, sample locations precomputed), fp32, TF32 off"}
    print(name, "reference ops on GPU %.3f ms | ours %.3f ms | x%.1f" % (t_ref, t_ours, t_ref / t_ours), flush=True)

H = W = 64; N = 4
P1, P2 = syn.pairs_from_ring(N, 4 * H)
tP1, tP2 = torch.from_numpy(P1.astype(np.float32)).to(dev), torch.from_numpy(P2.astype(np.float32)).to(dev)
for C in (64, 128, 256, 512):
    for K in (16, 32, 64, 128):
        f1 = torch.relu(torch.randn(N, C, H, W, device=dev)); f2 = torch.relu(torch.randn(N, C, H, W, device=dev))
        t = ev_time(lambda: epi.epipolar_fusion(f1, f2, tP1, tP2, K=K, correct_normalize=True), warm=5, iters=30)
        balg = 12 * N * C * H * W + 96 * N + 4 * N * K * H * W + 8 * N * H * W
        cfg = epi.make_cfg(KEYPOINT=dict(HEATMAP_SIZE=(H, W), NFEATS=C), EPIPOLAR=dict(SAMPLESIZE=K, USE_CORRECT_NORMALIZE=True))
        locs = None
        t_ref = None
        if K * C <= 64 * 256:
            from oracle import epipolar_oracle as eo
            locs = eo.sample_locs(cfg, P1.astype(np.float32), P2.astype(np.float32), H, W, K, dtype=np.float32, geometry="reference")
            t_ref = ev_time(lambda: torch_port.forward(cfg, f1, f2, P1, P2, locs=locs), warm=2, iters=5)
        rec = {"C": C, "K": K, "ms": t, "GBps": balg / t / 1e6, "frac": balg / t / 1e6 / PEAK, "kernel": "tile" if C <= 256 else "warp",
               "reference_ops_gpu_ms": t_ref}
        out["sweep"].append(rec)
        print(rec, flush=True)
json.dump(out, open("gpurun_out/sweep_r1.json", "w"), indent=1)
