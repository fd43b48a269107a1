"""-m gpu: the tcgen05 split-precision GEMM against an fp64 matmul."""
import pytest
import torch

from distar_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (300, 256, 256), (4096, 768, 256), (1000, 256, 1024),
                                   (129, 1024, 256), (2048, 1536, 1536), (5, 128, 384)])
@pytest.mark.parametrize('terms', [3, 1])
def test_gemm_against_fp64(M, N, K, terms):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    a_hi, a_lo = ops.split_bf16(a.to(DEV))
    w_hi, w_lo = ops.split_bf16(w.to(DEV))
    ref = a.double() @ w.double().t() + b.double()
    c = ops.gemm_split(a_hi, a_lo, w_hi, w_lo, b.to(DEV), relu=False, terms=terms)
    torch.cuda.synchronize()
    err = (c.double().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if terms == 3 else 2e-2
    assert err <= tol * scale, 'terms=%d M=%d N=%d K=%d: max err %.3e (scale %.3e)' % (terms, M, N, K, err, scale)
    # relu + split outputs
    c2, h, l = ops.gemm_split(a_hi, a_lo, w_hi, w_lo, b.to(DEV), relu=True, terms=terms, want_split=True)
    assert torch.equal(c2, torch.relu(c))
    assert (h.float() + l.float() - c2).abs().max().item() <= 1e-4 * max(scale, 1.0)


def test_linear_autograd_through_gemm():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 200, 256, generator=g)
    w = torch.randn(512, 256, generator=g) / 16
    b = torch.randn(512, generator=g)
    go = torch.randn(3, 200, 512, generator=g)
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    torch.relu(torch.nn.functional.linear(xr, wr, br)).backward(go)
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    y = ops.linear(xd, wd, bd, relu=True)
    y.backward(go.to(DEV))
    for got, ref, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        assert (got.cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item(), n
