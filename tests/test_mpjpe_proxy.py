"""North-star accuracy criterion on the synthetic MPJPE proxy: the fusion layer's effect on
triangulated 3-D joints must stay within 0.1 mm of what the reference's own layer gives on identical inputs.
The reference numbers were frozen in the build container (tests/golden/mpjpe_proxy.json)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, mpjpe_proxy as mp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mpjpe_proxy.json")))["seeds"]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_mpjpe_matches_reference(seed):
    d = mp.build(seed)
    o = c_oracle.forward(d["cfg"], d["feat_ref"], d["feat_src"], d["P_ref"], d["P_src"])
    got = mp.mpjpe(d, o["out"])
    assert abs(got - GOLD[str(seed)]["mpjpe_reference_mm"]) < 0.1
    assert abs(mp.mpjpe(d, np.zeros_like(d["feat_ref"])) - GOLD[str(seed)]["mpjpe_no_fusion_mm"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["auto", "warp"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cuda_mpjpe_within_0p1mm_of_reference(seed, variant):
    import torch
    import epipolar_transformers_b200 as epi
    d = mp.build(seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out, _, _, _ = epi.epipolar_fusion(t(d["feat_ref"]), t(d["feat_src"]), t(d["P_ref"]), t(d["P_src"]), K=mp.K,
                                       correct_normalize=True, variant=variant)
    got = mp.mpjpe(d, out.cpu().numpy())
    ref = GOLD[str(seed)]["mpjpe_reference_mm"]
    assert abs(got - ref) < 0.1, (got, ref)
