"""TEST INFRASTRUCTURE (not product code).

Runs the UNMODIFIED reference module /root/reference/modeling/layers/epipolar.py on CPU so
its outputs can be frozen as golden vectors (oracle/make_golden.py) and used to pin the
restatements in oracle/epipolar_oracle.py and oracle/epi_oracle.c.

yacs is not installed in this image; the reference only does attribute reads/writes on
its `cfg`, so `yacs.config.CfgNode` is shimmed by an attribute-dict (SURVEY.md 8c,
"verified").  /root/reference exists only in the build container: nothing that runs on
the GPU box may import this file.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("EPI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling", "layers", "epipolar.py"))


class _CN(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _install_shims():
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yacs_config = types.ModuleType("yacs.config")
        yacs_config.CfgNode = _CN
        yacs.config = yacs_config
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = yacs_config
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


_ref_mod = None


def load_reference():
    """Returns (epipolar_module, cfg) of the reference; imports once."""
    global _ref_mod
    if _ref_mod is None:
        if not reference_available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        _install_shims()
        warnings.filterwarnings("ignore")
        import importlib
        # import the leaf modules directly: modeling/__init__.py pulls the whole model zoo
        # (cv2, torchvision, ...) which the hot path does not need.
        core = importlib.import_module("core")
        pkg_modeling = types.ModuleType("modeling"); pkg_modeling.__path__ = [os.path.join(REFERENCE_ROOT, "modeling")]
        pkg_layers = types.ModuleType("modeling.layers"); pkg_layers.__path__ = [os.path.join(REFERENCE_ROOT, "modeling", "layers")]
        pkg_vision = types.ModuleType("vision"); pkg_vision.__path__ = [os.path.join(REFERENCE_ROOT, "vision")]
        sys.modules.setdefault("modeling", pkg_modeling)
        sys.modules.setdefault("modeling.layers", pkg_layers)
        sys.modules.setdefault("vision", pkg_vision)
        _ref_mod = (importlib.import_module("modeling.layers.epipolar"), core.cfg)
    return _ref_mod


def apply_cfg(ref_cfg, our_cfg):
    """Copy the hot-path keys of our duck-typed cfg onto the reference's global cfg."""
    ref_cfg.BACKBONE.DOWNSAMPLE = our_cfg.BACKBONE.DOWNSAMPLE
    ref_cfg.BACKBONE.BODY = our_cfg.BACKBONE.BODY
    ref_cfg.KEYPOINT.HEATMAP_SIZE = tuple(our_cfg.KEYPOINT.HEATMAP_SIZE)
    ref_cfg.KEYPOINT.NFEATS = our_cfg.KEYPOINT.NFEATS
    ref_cfg.DATASETS.IMAGE_RESIZE = our_cfg.DATASETS.IMAGE_RESIZE
    ref_cfg.DATASETS.PREDICT_RESIZE = our_cfg.DATASETS.PREDICT_RESIZE
    ref_cfg.DATASETS.CAMERAS = tuple(our_cfg.DATASETS.CAMERAS)
    for k, v in our_cfg.EPIPOLAR.items():
        ref_cfg.EPIPOLAR[k] = v
    ref_cfg.VIS.EPIPOLAR_LINE = our_cfg.VIS.EPIPOLAR_LINE


def run_reference(our_cfg, feat_ref, feat_src, P_ref, P_src, params=None, dtype_P="float32",
                  train_mode=False, threads=None):
    """Run the reference Epipolar.forward on CPU.

    feat_*: numpy float32 [N,C,H,W]; P_*: numpy [N,3,4] (cast to dtype_P, the reference
    casts float64->float32 in modeling/model.py:183-195; float64 P gives the fp64 geometry of
    the T2 protocol).  params: optional dict {'z.weight',...} loaded into the module.
    Returns dict(out, corr_pos, attn, sample_locs [K,N,H,W,2]) as numpy.
    """
    import numpy as np
    import torch
    mod, ref_cfg = load_reference()
    apply_cfg(ref_cfg, our_cfg)
    ref_cfg.VIS.EPIPOLAR_LINE = True          # makes forward return sample_locs (epipolar.py:266-267)
    if threads:
        torch.set_num_threads(threads)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = mod.Epipolar()
        if params:
            sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
            missing = m.load_state_dict(sd, strict=False)
            assert not missing.unexpected_keys, missing
        m.train(train_mode)
        tdt = getattr(torch, dtype_P)
        f1 = torch.from_numpy(np.ascontiguousarray(feat_ref))
        f2 = torch.from_numpy(np.ascontiguousarray(feat_src))
        P1 = torch.from_numpy(np.asarray(P_ref)).to(tdt)
        P2 = torch.from_numpy(np.asarray(P_src)).to(tdt)
        with torch.no_grad():
            out, corr_pos, attn, locs_t = m(f1, f2, P1, P2)
    return {
        "out": out.numpy(),
        "corr_pos": corr_pos.numpy(),
        "attn": attn.numpy(),
        "sample_locs": locs_t.transpose(0, 1).contiguous().numpy(),   # back to [K,N,H,W,2]
    }


def reference_sample_locs(our_cfg, P_ref, P_src, H, W, dtype="float64"):
    """grid2sample_locs alone (epipolar.py:323-418) at the requested precision -> [K,N,H,W,2]."""
    import numpy as np
    import torch
    mod, ref_cfg = load_reference()
    apply_cfg(ref_cfg, our_cfg)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = mod.Epipolar()
        tdt = getattr(torch, dtype)
        P1 = torch.from_numpy(np.asarray(P_ref)).to(tdt)
        P2 = torch.from_numpy(np.asarray(P_src)).to(tdt)
        m.sample_steps = m.sample_steps.to(tdt)
        with torch.no_grad():
            locs = m.grid2sample_locs(m.grid.to(tdt), P1, P2, H, W)
    return locs.numpy()
