"""-m gpu: implicit-GEMM convolutions (tcgen05) against fp64 F.conv2d."""
import pytest
import torch
import torch.nn.functional as F

from distar_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,relu,res', [
    (3, 16, 16, 128, 128, 3, True, False),
    (2, 16, 16, 128, 128, 3, True, True),
    (2, 32, 32, 64, 128, 3, True, False),
    (2, 64, 64, 32, 64, 3, True, False),      # 32 real input channels inside a 64-channel padded tensor
    (2, 32, 32, 128, 64, 3, False, False),
    (1, 64, 64, 64, 32, 3, True, False),      # 32 real output channels padded to 64
    (4, 16, 16, 128, 128, 1, True, False),
])
def test_conv_nhwc_fwd_bwd(N, H, W, Cin, Cout, k, relu, res):
    g = torch.Generator().manual_seed(H + Cin + Cout + k)
    cpad = (Cin + 63) // 64 * 64
    opad = (Cout + 63) // 64 * 64
    x = torch.zeros(N, H, W, cpad)
    x[..., :Cin] = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    r = torch.zeros(N, H, W, opad)
    r[..., :Cout] = torch.randn(N, H, W, Cout, generator=g)
    go = torch.zeros(N, H, W, opad)
    go[..., :Cout] = torch.randn(N, H, W, Cout, generator=g)
    xr, wr, br, rr = [t.double().clone().requires_grad_(True) for t in (x, w, b, r)]
    y = F.conv2d(xr[..., :Cin].permute(0, 3, 1, 2), wr, br, padding=k // 2).permute(0, 2, 3, 1)
    if res:
        y = y + rr[..., :Cout]
    pre = y
    if relu:
        y = torch.relu(y)
    # keep the comparison away from ReLU decision boundaries
    safe = (pre.detach().abs() > 1e-4) if relu else torch.ones_like(pre, dtype=torch.bool)
    (y * go[..., :Cout].double() * safe).sum().backward()
    xd, wd, bd, rd = [t.to(DEV).requires_grad_(True) for t in (x, w, b, r)]
    out = ops.conv_nhwc(xd, wd, bd, relu=relu, residual=rd if res else None)
    assert out.shape == (N, H, W, opad)
    gom = go.clone()
    gom[..., :Cout] = gom[..., :Cout] * safe
    (out * gom.to(DEV)).sum().backward()
    ref = y.detach()
    err = ((out[..., :Cout].double().cpu() - ref).abs() * safe).max().item()
    assert err <= 2e-5 * ref.abs().max().item(), ('fwd', err, ref.abs().max().item())
    assert out[..., Cout:].abs().max().item() == 0 if opad > Cout else True
    for got, want, name in [(xd.grad[..., :Cin], xr.grad[..., :Cin], 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        e = (got.double().cpu() - want).abs().max().item()
        assert e <= 1e-4 * want.abs().max().item(), (name, e, want.abs().max().item())
    if res:
        e = (rd.grad[..., :Cout].double().cpu() - rr.grad[..., :Cout]).abs().max().item()
        assert e <= 1e-4 * rr.grad.abs().max().item(), ('dres', e)


@pytest.mark.parametrize('N,H,W,Cin,C,pair_only', [(3, 16, 16, 128, 64, True), (2, 32, 32, 64, 32, False),
                                                   (1, 8, 16, 64, 32, False), (2, 8, 8, 128, 64, False)])
def test_upsample_conv3x3_fwd_bwd(N, H, W, Cin, C, pair_only):
    """act(conv3x3(bilinear2x(x))) through the low-resolution factorisation against fp64 interpolate + conv2d
    (head/action_arg_head.py:436-443): forward, dX, dW, db; borders (clamped interpolation, zero conv padding) included."""
    g = torch.Generator().manual_seed(H + Cin + C)
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(C, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(C, generator=g)
    go = torch.randn(N, 2 * H, 2 * W, C, generator=g)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    up = F.interpolate(xr.permute(0, 3, 1, 2), scale_factor=2., mode='bilinear')
    ref = torch.relu(F.conv2d(up, wr, br, padding=1)).permute(0, 2, 3, 1)
    # keep the comparison away from ReLU boundaries: kill the gradient where the pre-activation is within 1e-4 of 0
    pre = F.conv2d(up, wr, br, padding=1).permute(0, 2, 3, 1).detach()
    go = go * (pre.abs() > 1e-4).float()
    ref.backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    y = ops.upsample_conv3x3(xd, wd, bd, True, pair_only=pair_only)
    y.backward(go.to(DEV))
    if pair_only:
        hi, lo = y._dsb_split
        val = hi.double().cpu() + lo.double().cpu()
    else:
        val = y.detach().double().cpu()
    scale = ref.abs().max().item()
    assert (val - ref.detach()).abs().max().item() <= 3e-5 * scale
    for got, want, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        err = (got.double().cpu() - want).abs().max().item()
        assert err <= 1e-4 * want.abs().max().item(), (n, err, want.abs().max().item())
