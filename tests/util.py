"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def rel_max(a, b):
    """max|a-b| / max|b|  (the T1 metric of SURVEY.md section 8c)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def px_err(locs_a, locs_b, H, W, far=50.0):
    """Max sample-location error in feature pixels over samples that are not far-sentinels in b;
    also returns whether the far-sentinel sets agree."""
    a = np.asarray(locs_a, np.float64); b = np.asarray(locs_b, np.float64)
    fa = np.abs(a).max(-1) >= far
    fb = np.abs(b).max(-1) >= far
    scale = np.array([W / 2.0, H / 2.0])
    d = np.abs(a - b) * scale
    d = d[~fb & ~fa]
    return (float(d.max()) if d.size else 0.0), bool((fa == fb).all())
