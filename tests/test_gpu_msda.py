"""Multi-scale deformable attention kernels (csrc/msda.cu) against the op's grid_sample definition (oracle/msda.py):
forward within 1e-5, gradients within rtol 1e-3 of the float64 evaluation (the scatter into d value uses fp32 atomics)."""
import pytest
import torch

from dskd_b200.harness.msda import ms_deform_attn
from dskd_b200._lib import DskdError
from oracle.msda import msda_torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _case(N, Lq, M, D, shapes, P, seed=0, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g)
    # locations around [0,1] with some points outside the maps and some exactly on cell centres / borders
    loc = (torch.rand(N, Lq, M, len(shapes), P, 2, generator=g) - 0.5) * (1.0 + spread) + 0.5
    loc[0, 0, 0, 0, 0] = torch.tensor([0.0, 0.0])
    loc[0, 0, 0, 0, 1 % P] = torch.tensor([1.0, 1.0])
    loc[0, 0, 0, -1, 0] = torch.tensor([0.5 / shapes[-1][1], 0.5 / shapes[-1][0]])      # centre of cell (0, 0)
    attn = torch.rand(N, Lq, M, len(shapes) * P, generator=g).softmax(-1).view(N, Lq, M, len(shapes), P)
    return value, loc, attn


@pytest.mark.parametrize('cfg', [
    dict(N=2, Lq=50, M=8, D=32, shapes=[(13, 17), (7, 9), (4, 5), (2, 3)], P=4),            # the detector's geometry
    dict(N=1, Lq=333, M=4, D=32, shapes=[(25, 42), (13, 21)], P=4, spread=0.2),
    dict(N=2, Lq=7, M=2, D=16, shapes=[(5, 6), (3, 3), (2, 2)], P=2),                        # D < 32, L*P < 16
    dict(N=1, Lq=9, M=2, D=64, shapes=[(6, 5)], P=3, spread=3.0),                            # D > 32, far outside
], ids=['detr', 'two-level', 'small-head', 'wide-head'])
def test_forward_backward_vs_grid_sample_definition(cfg):
    shapes = cfg['shapes']
    value, loc, attn = _case(cfg['N'], cfg['Lq'], cfg['M'], cfg['D'], shapes, cfg['P'], spread=cfg.get('spread', 1.0))
    v, l, a = (t.clone().to(DEV).requires_grad_(True) for t in (value, loc, attn))
    out = ms_deform_attn(v, shapes, l, a)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(go.to(DEV))
    v64, l64, a64 = (t.double().requires_grad_(True) for t in (value, loc, attn))
    ref = msda_torch(v64, shapes, l64, a64)
    ref.backward(go.double())
    torch.testing.assert_close(out.detach().cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    for got, want in ((v.grad, v64.grad), (l.grad, l64.grad), (a.grad, a64.grad)):
        atol = 1e-5 * float(want.abs().max())
        torch.testing.assert_close(got.cpu().double(), want, rtol=1e-3, atol=atol)
    # the fp32 evaluation of the definition is as far from float64 as the kernel is
    v32, l32, a32 = (t.clone().requires_grad_(True) for t in (value, loc, attn))
    ref32 = msda_torch(v32, shapes, l32, a32)
    torch.testing.assert_close(out.detach().cpu(), ref32.detach(), rtol=1e-4, atol=1e-5)


def test_encoder_sized_call_and_errors():
    shapes = [(100, 167), (50, 84), (25, 42), (13, 21)]
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device=DEV).manual_seed(0)
    value = torch.randn(1, S, 8, 32, device=DEV, generator=g, requires_grad=True)
    loc = torch.rand(1, S, 8, 4, 4, 2, device=DEV, generator=g, requires_grad=True)
    attn = torch.rand(1, S, 8, 16, device=DEV, generator=g).softmax(-1).view(1, S, 8, 4, 4).requires_grad_(True)
    out = ms_deform_attn(value, shapes, loc, attn)
    out.sum().backward()
    # linearity in the weights: sum over channels and queries of out == sum of attn * sampled; with all-ones upstream
    # gradient d attn = sum_d sampled, hence <attn, d attn> == out.sum()
    torch.testing.assert_close((attn.detach() * attn.grad).sum(), out.detach().sum(), rtol=1e-4, atol=1e-2)
    # every token's gradient is a sum of non-negative bilinear weights times attn >= 0
    assert float(value.grad.min()) >= 0.0 and torch.isfinite(value.grad).all()
    with pytest.raises(DskdError):
        ms_deform_attn(value.detach().cpu(), shapes, loc.detach().cpu(), attn.detach().cpu())
    with pytest.raises(DskdError):
        ms_deform_attn(value.detach(), shapes[:3], loc.detach(), attn.detach())
