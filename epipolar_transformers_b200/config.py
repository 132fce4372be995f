"""Duck-typed configuration tree for the epipolar fusion path.

The reference module takes no constructor arguments: it reads a global yacs tree
(`from core import cfg`, /root/reference/core/__init__.py:1) whose epipolar keys are
declared at /root/reference/core/config.py:69-118.  This file provides the same
attribute paths with the same defaults so that `Epipolar()` here is constructed the
same way, without requiring yacs (not installed on the target image).  Any object with
the same attributes (e.g. the reference's real yacs `cfg`) can be passed instead via
`Epipolar(cfg=...)` or installed globally with `set_global_cfg`.

Only keys the hot path reads are present; see SURVEY.md section 5 ("Config / flags").
"""
from __future__ import annotations

import copy


class Node(dict):
    """Minimal attribute-dict (yacs.CfgNode look-alike: attribute reads + nested merge)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - error path
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), Node):
                self[k].merge(v)
            else:
                self[k] = Node(v) if isinstance(v, dict) and not isinstance(v, Node) else v
        return self


def default_cfg() -> Node:
    """Defaults copied *by value* from /root/reference/core/config.py (line in comment)."""
    c = Node()
    c.BACKBONE = Node(BODY="epipolarposeR-50", DOWNSAMPLE=4)            # :20,:23
    c.KEYPOINT = Node(HEATMAP_SIZE=(64, 64), NFEATS=256)                # :52 (224,224 default; 64 for 256^2 configs), :55
    c.DATASETS = Node(IMAGE_RESIZE=1.0, PREDICT_RESIZE=1.0, CAMERAS=())  # :168,:170 (YAMLs under configs/epipolar set 1.)
    c.EPIPOLAR = Node(
        ATTENTION="avg",              # :81 default 'max'; every shipped epipolar YAML sets avg
        SIMILARITY="dot",             # :82
        SAMPLESIZE=64,                # :84
        SOFTMAX_ENABLED=True,         # :85
        SOFTMAXSCALE=1.0 / 64 ** 0.5,  # :86 evaluated once from the *default* SAMPLESIZE
        MERGE="late",                 # :89
        OTHER_GRAD=("other1", "other2"),  # :93
        SHARE_WEIGHTS=False,          # :95
        PARAMETERIZED=(),             # :98
        ZRESIDUAL=False,              # :99
        MULTITEST=False,              # :101
        PRIOR=False,                  # :104
        PRIORMUL=False,               # :105
        REPROJECT_LOSS_WEIGHT=0.0,    # :107
        FIND_CORR="feature",          # :113
        BOTTLENECK=1,                 # :115
        POOLING=False,                # :116
        USE_CORRECT_NORMALIZE=False,  # :118
    )
    c.VIS = Node(EPIPOLAR_LINE=False)                                   # :290
    return c


_GLOBAL = default_cfg()


def get_global_cfg() -> Node:
    return _GLOBAL


def set_global_cfg(c) -> None:
    """Install a cfg (ours or the reference's yacs node) as the module-level default."""
    global _GLOBAL
    _GLOBAL = c


def make_cfg(**overrides) -> Node:
    """default_cfg() + nested overrides, e.g. make_cfg(EPIPOLAR=dict(SAMPLESIZE=32))."""
    c = default_cfg()
    c.merge(overrides)
    return c


# Named shapes used by BASELINE.json configs.
def cfg_h36m_r50_256() -> Node:
    """configs/epipolar/keypoint_h36m_zresidual_fixed.yaml:27-39 (config 2)."""
    return make_cfg(
        KEYPOINT=dict(HEATMAP_SIZE=(64, 64), NFEATS=256),
        EPIPOLAR=dict(PARAMETERIZED=("z",), ZRESIDUAL=True, USE_CORRECT_NORMALIZE=True),
    )


def cfg_h36m_r152_384() -> Node:
    """configs/epipolar/keypoint_h36m_resnet152_384.yaml:25-33 (config 3)."""
    return make_cfg(KEYPOINT=dict(HEATMAP_SIZE=(96, 96), NFEATS=256))
