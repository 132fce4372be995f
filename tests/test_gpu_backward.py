"""GPU: backward of the fused attention against PyTorch autograd through the reference's own op
composition (F.grid_sample x2, mul/sum, ==0 mask assignment, softmax, weighted sum — epipolar.py:188-247) evaluated in
float64 on the same sample locations.  Tolerance 1e-4 relative to max|grad| (fp32 kernel, float atomics)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import epipolar_transformers_b200 as epi
from oracle import golden_cases as gc
from tests.util import rel_max

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def torch_reference(f1, f2k, f2v, locs, scale):
    """differentiable restatement in the dtype of the inputs; f2k / f2v = the source map as 'other1' / 'other2'."""
    N, C, H, W = f1.shape
    K = locs.shape[0]
    outs, attns = [], []
    for n in range(N):
        g = locs[:, n]
        keys = F.grid_sample(f2k[n].unsqueeze(0).expand(K, -1, -1, -1), g, align_corners=False)
        vals = F.grid_sample(f2v[n].unsqueeze(0).expand(K, -1, -1, -1), g, align_corners=False)
        sim = (keys * f1[n].unsqueeze(0)).sum(1)
        sim = torch.where(sim == 0, torch.full_like(sim, -1e10), sim)          # `sim[sim==0] = -1e10`: no gradient there
        a = F.softmax(sim * scale, 0)
        outs.append((vals * a.unsqueeze(1)).sum(0))
        attns.append(a)
    return torch.stack(outs), torch.stack(attns)


@pytest.mark.parametrize("name", ["tiny_ring_z", "tiny_randn_krt", "tiny_zero_query", "cfg1_ring"])
@pytest.mark.parametrize("other_grad", [("other1", "other2"), ("other2",), ("other1",)])
def test_backward_vs_autograd_fp64(name, other_grad):
    cfg, f1, f2, P1, P2, _ = gc.build_inputs(name)
    spec = gc.CASES[name]
    if name == "cfg1_ring" and other_grad != ("other1", "other2"):
        pytest.skip("one OTHER_GRAD setting is enough at this size")
    K = spec["K"]
    t1 = dev(f1).requires_grad_(True); t2 = dev(f2).requires_grad_(True)
    opts = dict(fwd=dict(K=K, downsample=cfg.BACKBONE.DOWNSAMPLE, img_scale=cfg.DATASETS.IMAGE_RESIZE * cfg.DATASETS.PREDICT_RESIZE,
                         softmax_scale=cfg.EPIPOLAR.SOFTMAXSCALE, correct_normalize=spec["correct"], align_corners=False,
                         want_corr=True, want_locs=True, variant="auto"),
                grad_keys="other1" in other_grad, grad_vals="other2" in other_grad)
    from epipolar_transformers_b200.epipolar import _FusionFn
    out, corr, attn, locs = _FusionFn.apply(t1, t2, dev(P1), dev(P2), opts)
    torch.manual_seed(3)
    w_out = torch.randn_like(out); w_attn = torch.randn_like(attn)
    loss = (out * w_out).sum() + 0.3 * (attn * w_attn).sum()                    # a loss that also touches the attention output
    g1, g2 = torch.autograd.grad(loss, (t1, t2))
    # reference in float64 on the locations the kernel sampled
    r1 = dev(f1).double().requires_grad_(True); r2 = dev(f2).double().requires_grad_(True)
    r2k = r2 if "other1" in other_grad else r2.detach()
    r2v = r2 if "other2" in other_grad else r2.detach()
    ro, ra = torch_reference(r1, r2k, r2v, locs.double(), float(cfg.EPIPOLAR.SOFTMAXSCALE))
    rloss = (ro * w_out.double()).sum() + 0.3 * (ra * w_attn.double()).sum()
    e1, e2 = torch.autograd.grad(rloss, (r1, r2), allow_unused=True)
    assert rel_max(out.detach().cpu().numpy(), ro.detach().cpu().numpy()) < 1e-4
    assert rel_max(g1.cpu().numpy(), e1.cpu().numpy()) < 1e-4
    assert rel_max(g2.cpu().numpy(), e2.cpu().numpy()) < 1e-4


def test_module_trains_end_to_end():
    """Epipolar under autograd in train mode (engine/trainer.py:72): gradients reach both feature maps and z / bn, and
    match the PyTorch composition (conv1x1 + BatchNorm(train) + ZRESIDUAL on top of the reference attention)."""
    name = "tiny_ring_z"
    cfg, f1, f2, P1, P2, params = gc.build_inputs(name)
    cfg.VIS.EPIPOLAR_LINE = True
    torch.backends.cudnn.allow_tf32 = False            # the module's own conv1x1 (PyTorch) must not run in TF32 for a 1e-4 comparison
    torch.backends.cuda.matmul.allow_tf32 = False
    m = epi.Epipolar(cfg=cfg).cuda().train()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    t1 = dev(f1).requires_grad_(True); t2 = dev(f2).requires_grad_(True)
    out, corr, attn, locs_t = m(t1, t2, dev(P1), dev(P2))
    w = torch.randn_like(out)
    (out * w).sum().backward()
    assert t1.grad is not None and t2.grad is not None and m.z.weight.grad is not None and m.bn.weight.grad is not None
    # reference composition in float64
    mz = torch.nn.Conv2d(m.z.in_channels, m.z.out_channels, 1).cuda().double()
    mz.load_state_dict({k: v.double() for k, v in m.z.state_dict().items()})
    r1 = dev(f1).double().requires_grad_(True); r2 = dev(f2).double().requires_grad_(True)
    locs = locs_t.transpose(0, 1).contiguous().double()
    ro, _ = torch_reference(r1, r2, r2, locs, float(cfg.EPIPOLAR.SOFTMAXSCALE))
    y = F.batch_norm(mz(ro), None, None, m.bn.weight.double(), m.bn.bias.double(), True, 0.1, m.bn.eps) + ro
    (y * w.double()).sum().backward()
    assert rel_max(out.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-4
    assert rel_max(t1.grad.cpu().numpy(), r1.grad.cpu().numpy()) < 2e-4
    assert rel_max(t2.grad.cpu().numpy(), r2.grad.cpu().numpy()) < 2e-4
    assert rel_max(m.z.weight.grad.cpu().numpy(), mz.weight.grad.cpu().numpy()) < 2e-4
