"""CPU: pins both oracle restatements against golden vectors frozen from the reference itself
(oracle/make_golden.py).  T1 = given the reference's own sample locations, everything after the
geometry must agree to fp32 round-off; T2 = oracle geometry vs the reference run in fp64."""
import numpy as np
import pytest

from oracle import c_oracle, epipolar_oracle as eo, golden_cases as gc
from tests.util import load_golden, px_err, rel_max

FULL = [n for n, s in gc.CASES.items() if s["full"]]
BIG = [n for n, s in gc.CASES.items() if not s["full"]]


@pytest.mark.parametrize("name", FULL)
def test_numpy_oracle_T1_full(name):
    g = load_golden(name)
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    o = eo.forward(cfg, f1, f2, P1, P2, params=params, locs=g["sample_locs"])
    assert rel_max(o["out"], g["out"]) < 2e-6
    assert rel_max(o["attn"], g["attn"]) < 2e-6
    # argmax ties between equal softmax weights may legitimately pick another sample
    same = np.abs(o["corr_pos"] - g["corr_pos"]).max(-1) < 1e-4
    assert same.mean() > 0.995


@pytest.mark.parametrize("name", FULL)
def test_c_oracle_T1_full(name):
    g = load_golden(name)
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    o = c_oracle.forward(cfg, f1, f2, P1, P2, locs=g["sample_locs"])
    out = eo.z_epilogue(o["out"], params, cfg.EPIPOLAR.ZRESIDUAL) if params else o["out"]
    assert rel_max(out, g["out"]) < 2e-6
    assert rel_max(o["attn"], g["attn"]) < 2e-6
    same = np.abs(o["corr_pos"] - g["corr_pos"]).max(-1) < 1e-4
    assert same.mean() > 0.995


@pytest.mark.parametrize("name", FULL)
@pytest.mark.parametrize("geometry", ["reference", "hinf"])
def test_numpy_geometry_T2(name, geometry):
    """fp64 restatement (both line parameterisations) vs the reference executed in fp64."""
    g = load_golden(name)
    spec = gc.CASES[name]
    cfg, _, _, P1, P2, _ = gc.build_inputs(name)
    locs = eo.sample_locs(cfg, P1, P2, spec["H"], spec["W"], dtype=np.float64, geometry=geometry)
    err, far_ok = px_err(locs, g["sample_locs_fp64"], spec["H"], spec["W"])
    assert far_ok
    assert err < 1e-5


@pytest.mark.parametrize("name", FULL + BIG)
def test_c_geometry_T2(name):
    """C oracle geometry (fp64 and fp32 H-infinity form) vs the reference in fp64; the fp32
    form must be at least as close to fp64 truth as the reference's own fp32 locations."""
    g = load_golden(name)
    spec = gc.CASES[name]
    H, W = spec["H"], spec["W"]
    cfg, f1, f2, P1, P2, _ = gc.build_inputs(name)
    for fp32, tol in ((False, 1e-5), (True, 2e-3)):
        o = c_oracle.forward(cfg, f1[:, :1], f2[:, :1], P1, P2, geom_fp32=fp32)
        locs = o["sample_locs"]
        if not spec["full"]:
            px = g["pixels"]; n_idx = np.arange(spec["N"])[:, None]
            locs = locs.transpose(1, 2, 3, 0, 4)[n_idx, px[..., 0], px[..., 1]]
        err, far_ok = px_err(locs, g["sample_locs_fp64"], H, W)
        ref_err, _ = px_err(g["sample_locs"], g["sample_locs_fp64"], H, W)
        assert far_ok
        assert err < tol, (err, ref_err)
        if fp32 and spec["cams"] != "randn":
            assert err <= max(ref_err, 1e-4), (err, ref_err)


@pytest.mark.parametrize("name", BIG)
def test_c_oracle_T1_subsampled(name):
    """BASELINE-sized cases: inject the reference's frozen sample locations at the frozen pixels."""
    g = load_golden(name)
    spec = gc.CASES[name]
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    base = c_oracle.forward(cfg, f1[:, :1], f2[:, :1], P1, P2)["sample_locs"]      # own geometry everywhere
    px = g["pixels"]; n_idx = np.arange(spec["N"])[:, None]
    lv = base.transpose(1, 2, 3, 0, 4)          # view [N,H,W,K,2]
    lv[n_idx, px[..., 0], px[..., 1]] = g["sample_locs"]
    o = c_oracle.forward(cfg, f1, f2, P1, P2, locs=base)
    out = eo.z_epilogue(o["out"], params, cfg.EPIPOLAR.ZRESIDUAL) if params else o["out"]
    got_out = out[n_idx, :, px[..., 0], px[..., 1]]
    got_attn = o["attn"][n_idx, :, px[..., 0], px[..., 1]]
    assert rel_max(got_out, g["out"]) < 1e-5
    assert rel_max(got_attn, g["attn"]) < 1e-5


@pytest.mark.parametrize("name", FULL)
def test_torch_port_T1_full(name):
    """the CPU-baseline port (same ATen op sequence as the reference) reproduces the golden vectors."""
    import torch
    from oracle import torch_port
    g = load_golden(name)
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    out, corr, attn = torch_port.forward(cfg, torch.from_numpy(f1), torch.from_numpy(f2), P1, P2, params=params,
                                         locs=g["sample_locs"])
    assert rel_max(out.numpy(), g["out"]) < 2e-6
    assert rel_max(attn.numpy(), g["attn"]) < 2e-6
    assert (np.abs(corr.numpy() - g["corr_pos"]).max(-1) < 1e-4).mean() > 0.995


def test_torch_port_own_geometry_close_to_reference():
    import torch
    from oracle import torch_port
    name = "tiny_ring_z"
    g = load_golden(name)
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    out, _, attn = torch_port.forward(cfg, torch.from_numpy(f1), torch.from_numpy(f2), P1, P2, params=params)
    assert rel_max(attn.numpy(), g["attn"]) < 5e-2       # fp32 pinv geometry noise, not a parity bar
