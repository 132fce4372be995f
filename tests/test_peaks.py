"""find_tensor_peak_batch: the numpy oracle is pinned to vectors frozen from the reference
function (CPU test); the CUDA kernel is compared with those vectors and the oracle through the C ABI (GPU test)."""
import os

import numpy as np
import pytest

from oracle import make_golden_peaks as mg
from oracle import peaks_oracle as po

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peaks.npz"))


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_matches_reference_golden(name):
    J, H, W, radius, ds, _ = mg.CASES[name]
    locs, score = po.find_tensor_peak_batch(mg.heatmaps(name), radius, ds)
    np.testing.assert_allclose(locs, GOLD[name + "_locs"], rtol=0, atol=2e-4)       # image px (a few fp32 ulp of ~250)
    np.testing.assert_array_equal(score, GOLD[name + "_score"])


def test_oracle_integer_division_variant():
    """torch < 1.4 semantics (integer `index / W`): y moves by the dropped fraction times the stride."""
    name = "h36m_r50"
    J, H, W, radius, ds, _ = mg.CASES[name]
    a, _ = po.find_tensor_peak_batch(mg.heatmaps(name), radius, ds, int_div=False)
    b, _ = po.find_tensor_peak_batch(mg.heatmaps(name), radius, ds, int_div=True)
    assert np.abs(a[:, 0] - b[:, 0]).max() < 0.6 * ds and np.abs(a[:, 1] - b[:, 1]).max() < 1.5 * ds
    assert np.abs(a - b).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_kernel_matches_reference_golden(name):
    import torch
    import epipolar_transformers_b200 as epi
    J, H, W, radius, ds, _ = mg.CASES[name]
    h = torch.from_numpy(mg.heatmaps(name)).cuda()
    locs, score = epi.find_tensor_peak_batch(h, radius, ds)
    assert locs.shape == (J, 2) and score.shape == (J,)
    np.testing.assert_allclose(locs.cpu().numpy(), GOLD[name + "_locs"], rtol=0, atol=3e-4)
    np.testing.assert_array_equal(score.cpu().numpy(), GOLD[name + "_score"])
    # batched form == per-item form (what the caller's loop at resnet.py:423-428 computes), bit exact
    hb = torch.stack([h, h.flip(0), h * 0.5])
    lb, sb = epi.find_tensor_peak_batch(hb, radius, ds)
    for i in range(3):
        li, si = epi.find_tensor_peak_batch(hb[i], radius, ds)
        assert torch.equal(lb[i], li) and torch.equal(sb[i], si)
    # integer-division variant against the oracle
    li, _ = epi.find_tensor_peak_batch(h, radius, ds, int_div=True)
    lo, _ = po.find_tensor_peak_batch(mg.heatmaps(name), radius, ds, int_div=True)
    np.testing.assert_allclose(li.cpu().numpy(), lo, rtol=0, atol=3e-4)
