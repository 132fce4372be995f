"""CPU: host-side mirror of the reference interface (no GPU needed): constructor contract, parameter names,
unsupported-flag behaviour, config helpers."""
import pytest
import torch

import epipolar_transformers_b200 as epi


def test_state_dict_names_match_reference():
    """names the reference's checkpoints carry for this module (modeling/layers/epipolar.py:64-65, BN.py:28-36)"""
    m = epi.Epipolar(cfg=epi.cfg_h36m_r50_256())
    assert sorted(m.state_dict().keys()) == sorted(
        ["z.weight", "z.bias", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var", "bn.num_batches_tracked"])
    assert tuple(m.z.weight.shape) == (256, 256, 1, 1)
    assert float(m.bn.weight.abs().sum()) == 0.0 and float(m.bn.bias.abs().sum()) == 0.0     # zero-init BN (BN.py:48-52)
    assert list(epi.Epipolar(cfg=epi.cfg_h36m_r152_384()).state_dict().keys()) == []            # PARAMETERIZED=()


@pytest.mark.parametrize("override", [
    dict(ATTENTION="max"), dict(SIMILARITY="cos"), dict(SIMILARITY="prior"), dict(PRIOR=True), dict(POOLING=True),
    dict(FIND_CORR="rgb"), dict(REPROJECT_LOSS_WEIGHT=1.0), dict(SOFTMAX_ENABLED=False),
    dict(PARAMETERIZED=("z", "theta", "phi", "g"), BOTTLENECK=2)])
def test_unsupported_flags_raise_at_construction(override):
    with pytest.raises(NotImplementedError):
        epi.Epipolar(cfg=epi.make_cfg(EPIPOLAR=override))
    with pytest.raises(NotImplementedError):
        epi.Epipolar(debug=True, cfg=epi.make_cfg())


def test_global_cfg_like_reference():
    """`Epipolar()` with no arguments reads the module-level cfg, like `from core import cfg` in the reference."""
    old = epi.get_global_cfg()
    try:
        epi.set_global_cfg(epi.make_cfg(KEYPOINT=dict(HEATMAP_SIZE=(96, 96)), EPIPOLAR=dict(SAMPLESIZE=85)))
        m = epi.Epipolar()
        assert (m.feat_h, m.feat_w, m.sample_size) == (96, 96, 85)
        assert abs(m.cfg.EPIPOLAR.SOFTMAXSCALE - 0.125) < 1e-12        # stays 1/sqrt(64) whatever SAMPLESIZE is
    finally:
        epi.set_global_cfg(old)


def test_cpu_tensors_are_rejected_not_silently_computed():
    m = epi.Epipolar(cfg=epi.make_cfg(KEYPOINT=dict(HEATMAP_SIZE=(8, 8), NFEATS=8), EPIPOLAR=dict(SAMPLESIZE=8)))
    x = torch.zeros(1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        m(x, x, torch.zeros(1, 3, 4), torch.zeros(1, 3, 4))
    with pytest.raises(NotImplementedError):
        m(x, x, torch.zeros(1, 3, 4), torch.zeros(1, 3, 4), depth=x)


def test_multiview_helpers_roundtrip():
    from epipolar_transformers_b200 import multiview as mv
    pts = torch.tensor([[0.0, 0.0], [63.0, 63.0], [10.5, 20.25]])
    for correct in (True, False):
        g = mv.normalize(pts, 64, 64, correct)
        assert torch.allclose(mv.de_normalize(g, 64, 64, correct), pts, atol=1e-5)
    assert mv.coord2pix(mv.pix2coord(7.0, 4), 4) == 7.0
