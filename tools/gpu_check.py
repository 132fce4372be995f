"""Quick on-GPU comparison of the kernel variants against the C oracle (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import epipolar_transformers_b200 as epi
from oracle import c_oracle, golden_cases as gc

def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()

names = sys.argv[1:] or ["tiny_ring_z", "tiny_randn_krt", "tiny_relu_znores", "tiny_k85", "tiny_ds8_resize", "tiny_zero_query",
                         "cfg1_ring", "cfg1_randn_krt", "cfg2_r50_256_randn", "cfg3_r152_384"]
for name in names:
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    spec = gc.CASES[name]
    res = {}
    for variant in ("warp", "pipe", "sector"):
        try:
            out, corr, attn, locs = epi.epipolar_fusion(dev(f1), dev(f2), dev(P1), dev(P2), K=spec["K"], downsample=cfg.BACKBONE.DOWNSAMPLE,
                img_scale=cfg.DATASETS.IMAGE_RESIZE * cfg.DATASETS.PREDICT_RESIZE, correct_normalize=spec["correct"], want_locs=True, variant=variant)
            torch.cuda.synchronize()
            res[variant] = (out.cpu().numpy(), attn.cpu().numpy(), corr.cpu().numpy(), locs.cpu().numpy())
        except Exception as e:
            print(name, variant, "FAILED", repr(e)[:300]); res[variant] = None
    if res["warp"] is None: continue
    o = c_oracle.forward(cfg, f1, f2, P1, P2, locs=res["warp"][3])
    for variant in ("warp", "pipe", "sector"):
        if res[variant] is None: continue
        out, attn, corr, locs = res[variant]
        eo_ = np.abs(out - o["out"]).max() / max(np.abs(o["out"]).max(), 1e-30)
        ea = np.abs(attn - o["attn"]).max()
        ec = (np.abs(corr - o["corr_pos"]).max(-1) < 1e-3).mean()
        el = np.abs(locs - res["warp"][3]).max()
        print("%-20s %-5s out %.2e attn %.2e corr_agree %.4f locs_diff %.1e nan=%d" % (name, variant, eo_, ea, ec, el, int(np.isnan(out).sum())))
