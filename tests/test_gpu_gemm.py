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
    # the pair must be exactly the split of the fp32 result, with or without the fp32 store
    eh, el = ops.split_bf16(c2.clone())
    assert torch.equal(h, eh) and torch.equal(l, el)
    c3, h3, l3 = ops.gemm_split(a_hi, a_lo, w_hi, w_lo, b.to(DEV), relu=True, terms=terms, want_split='only')
    assert torch.equal(h3, eh) and torch.equal(l3, el)
    assert c3.shape == c2.shape and bool(torch.isnan(c3).all())       # placeholder, never to be read


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


# ------------------------------------------------------------------------------------------ generalised GEMM (dsb_gemm_ex)
from distar_b200 import lib as _lib


def _split(x):
    return ops.split_bf16(x.to(DEV).contiguous())


def _check(c, ref, tol=2e-5):
    err = (c.double().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale, 'max err %.3e (scale %.3e)' % (err, scale)


@pytest.mark.parametrize('M,N,K', [(256, 128, 64), (384, 256, 512)])
def test_gemm_ex_b_mn_major(M, N, K):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g)
    bm = torch.randn(K, N, generator=g) / K ** 0.5          # B stored [K, N]: reduction index strided
    a_hi, a_lo = _split(a)
    b_hi, b_lo = _split(bm)
    c = torch.empty(M, N, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=b_hi, b_lo=b_lo, b_mn=1, alpha=1.0, terms=3, c=c, m=M, n=N, k=K,
                 batch=1, inner=1, splits=1)
    torch.cuda.synchronize()
    _check(c, a.double() @ bm.double())


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (256, 384, 256)])
def test_gemm_ex_a_mn_major(M, N, K):
    g = torch.Generator().manual_seed(2)
    am = torch.randn(K, M, generator=g)                      # A stored [K, M]
    w = torch.randn(N, K, generator=g) / K ** 0.5
    a_hi, a_lo = _split(am)
    b_hi, b_lo = _split(w)
    c = torch.empty(M, N, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=b_hi, b_lo=b_lo, a_mn=1, alpha=1.0, terms=3, c=c, m=M, n=N, k=K,
                 batch=1, inner=1, splits=1)
    torch.cuda.synchronize()
    _check(c, am.double().t() @ w.double().t())


@pytest.mark.parametrize('tokens,Nw,Kw,splits', [(512, 128, 128, 1), (4096, 256, 1024, 16), (8192, 768, 256, 32)])
def test_gemm_ex_weight_gradient_split_k(tokens, Nw, Kw, splits):
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(tokens, Nw, generator=g)
    x = torch.randn(tokens, Kw, generator=g)
    a_hi, a_lo = _split(dy)
    b_hi, b_lo = _split(x)
    part = torch.empty(splits * Nw, Kw, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=b_hi, b_lo=b_lo, a_mn=1, b_mn=1, alpha=1.0, terms=3, c=part, m=Nw, n=Kw,
                 k=tokens, batch=1, inner=1, splits=splits, c_row_split=Nw)
    torch.cuda.synchronize()
    dw = part.view(splits, Nw, Kw).sum(0)
    _check(dw, dy.double().t() @ x.double(), tol=5e-5)


def test_gemm_ex_batched_attention_products():
    obs, S, H, D = 3, 512, 2, 128
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(obs * S, 3 * H * D, generator=g)
    q_hi, q_lo = _split(qkv)
    # scores[o,h] = Q K^T / sqrt(D)
    sc = torch.empty(obs * H * S, S, device=DEV)
    _lib.gemm_ex(a_hi=q_hi, a_lo=q_lo, b_hi=q_hi, b_lo=q_lo, alpha=1.0 / D ** 0.5, terms=3, c=sc, m=S, n=S, k=D,
                 batch=obs * H, inner=H, splits=1, a_col_base=0, a_col_inner=D, a_row_outer=S,
                 b_col_base=H * D, b_col_inner=D, b_row_outer=S, c_row_outer=H * S, c_row_inner=S)
    torch.cuda.synchronize()
    x = qkv.double().view(obs, S, 3, H, D)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    ref = torch.einsum('oqhd,okhd->ohqk', q, k) / D ** 0.5
    _check(sc.view(obs, H, S, S), ref)
    # context[o,h] = P V   (V is MN-major inside the same activation)
    p = torch.softmax(ref, -1).float().reshape(obs * H * S, S)
    p_hi, p_lo = _split(p)
    ctx = torch.empty(obs * S, H * D, device=DEV)     # heads concatenated along the feature axis, as the reference
    _lib.gemm_ex(a_hi=p_hi, a_lo=p_lo, b_hi=q_hi, b_lo=q_lo, b_mn=1, alpha=1.0, terms=3, c=ctx, m=S, n=D, k=S,
                 batch=obs * H, inner=H, splits=1, a_row_outer=H * S, a_row_inner=S,
                 b_col_base=2 * H * D, b_col_inner=D, b_row_outer=S, c_row_outer=S, c_col_inner=D)
    torch.cuda.synchronize()
    ref_ctx = torch.einsum('ohqk,okhd->oqhd', p.double().view(obs, H, S, S), v)
    _check(ctx.view(obs, S, H, D), ref_ctx)


def test_linear_autograd_weight_grad_split_k():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 512, 256, generator=g)
    w = torch.randn(1024, 256, generator=g) / 16
    b = torch.randn(1024, generator=g)
    go = torch.randn(8, 512, 1024, generator=g)
    # no ReLU here: a pre-activation within rounding of 0 flips the mask and moves dx by a whole term
    xr, wr, br = [t.double().clone().requires_grad_(True) for t in (x, w, b)]
    torch.nn.functional.linear(xr, wr, br).backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    ops.linear(xd, wd, bd, relu=False).backward(go.to(DEV))
    for got, ref, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        err = (got.double().cpu() - ref).abs().max().item()
        assert err <= 1e-4 * ref.abs().max().item(), (n, err, ref.abs().max().item())


@pytest.mark.parametrize('entity_num', [[512, 512, 512], [512, 77, 1], [300, 0, 128]])
def test_entity_attention_fwd_bwd(entity_num):
    n, S, H, D = len(entity_num), 512, 2, 128
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(n, S, 3 * H * D, generator=g)
    go = torch.randn(n, S, H * D, generator=g)
    en = torch.tensor(entity_num)
    qr = qkv.double().clone().requires_grad_(True)
    q, k, v = qr.view(n, S, 3, H, D).permute(2, 0, 3, 1, 4)
    sc = torch.matmul(q, k.transpose(2, 3)) / D ** 0.5
    mask = torch.arange(S).unsqueeze(0) < en.unsqueeze(1)
    sc = sc.masked_fill(~mask.view(n, 1, 1, S), -1e9)
    ref = torch.matmul(torch.softmax(sc, -1), v).permute(0, 2, 1, 3).reshape(n, S, H * D)
    ref.backward(go.double())
    qd = qkv.to(DEV).requires_grad_(True)
    out = ops.entity_attention(qd, en.to(DEV), H, D)
    out.backward(go.to(DEV))
    # the context exists only as the bf16 (hi, lo) pair its consumer (the projection GEMM) reads
    hi, lo = out._dsb_split
    val = hi.double().cpu() + lo.double().cpu()
    err = (val - ref.detach()).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), ('fwd', err)
    gerr = (qd.grad.double().cpu() - qr.grad).abs().max().item()
    assert gerr <= 1e-4 * qr.grad.abs().max().item(), ('bwd', gerr, qr.grad.abs().max().item())


# ------------------------------------------------------------ 2-CTA clusters: TMA multicast (mc=2), cta_group::2 MMA pair (mc=4)
@pytest.mark.parametrize('mc', [2, 4])
@pytest.mark.parametrize('M,N,K,bn', [(4096, 256, 256, 128), (4224, 1024, 256, 256), (1280, 256, 1024, 256), (3 * 128, 512, 128, 128),
                                      (1000, 512, 192, 256)])
def test_gemm_ex_multicast_k_major(M, N, K, bn, mc):
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    c = torch.empty(M, N, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b.to(DEV), alpha=1.0, terms=3, c=c, m=M, n=N, k=K,
                 batch=1, inner=1, splits=1, bn=bn, mc=mc)
    torch.cuda.synchronize()
    _check(c, a.double() @ w.double().t() + b.double())
    # 1-term product and the bf16-pair epilogue through the same cluster mode
    c_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    c_lo = torch.empty_like(c_hi)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b.to(DEV), alpha=1.0, relu=1, terms=3, c=None, c_hi=c_hi,
                 c_lo=c_lo, m=M, n=N, k=K, batch=1, inner=1, splits=1, bn=bn, mc=mc)
    eh, el = ops.split_bf16(torch.relu(c))
    assert torch.equal(c_hi, eh) and torch.equal(c_lo, el)


@pytest.mark.parametrize('mc', [2, 4])
def test_gemm_ex_multicast_mn_major_and_split_k(mc):
    tokens, Nw, Kw, splits = 8192, 768, 256, 8
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(tokens, Nw, generator=g)
    x = torch.randn(tokens, Kw, generator=g)
    a_hi, a_lo = _split(dy)
    b_hi, b_lo = _split(x)
    gw = torch.zeros(Nw, Kw, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=b_hi, b_lo=b_lo, a_mn=1, b_mn=1, alpha=1.0, terms=3, c=gw, m=Nw, n=Kw,
                 k=tokens, batch=1, inner=1, splits=splits, c_accumulate=1, mc=mc)
    torch.cuda.synchronize()
    _check(gw, dy.double().t() @ x.double(), tol=5e-5)


@pytest.mark.parametrize('mc', [2, 4])
def test_gemm_ex_multicast_batched_attention(mc):
    obs, S, H, D = 5, 512, 2, 128
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(obs * S, 3 * H * D, generator=g)
    q_hi, q_lo = _split(qkv)
    sc = torch.empty(obs * H * S, S, device=DEV)
    _lib.gemm_ex(a_hi=q_hi, a_lo=q_lo, b_hi=q_hi, b_lo=q_lo, alpha=1.0 / D ** 0.5, terms=3, c=sc, m=S, n=S, k=D,
                 batch=obs * H, inner=H, splits=1, a_col_base=0, a_col_inner=D, a_row_outer=S,
                 b_col_base=H * D, b_col_inner=D, b_row_outer=S, c_row_outer=H * S, c_row_inner=S, mc=mc)
    torch.cuda.synchronize()
    x = qkv.double().view(obs, S, 3, H, D)
    ref = torch.einsum('oqhd,okhd->ohqk', x[:, :, 0], x[:, :, 1]) / D ** 0.5
    _check(sc.view(obs, H, S, S), ref)
    p = torch.softmax(ref, -1).float().reshape(obs * H * S, S)
    p_hi, p_lo = _split(p)
    ctx = torch.empty(obs * S, H * D, device=DEV)
    _lib.gemm_ex(a_hi=p_hi, a_lo=p_lo, b_hi=q_hi, b_lo=q_lo, b_mn=1, alpha=1.0, terms=3, c=ctx, m=S, n=D, k=S,
                 batch=obs * H, inner=H, splits=1, a_row_outer=H * S, a_row_inner=S,
                 b_col_base=2 * H * D, b_col_inner=D, b_row_outer=S, c_row_outer=S, c_col_inner=D, mc=mc)
    torch.cuda.synchronize()
    _check(ctx.view(obs, S, H, D), torch.einsum('ohqk,okhd->oqhd', p.double().view(obs, H, S, S), x[:, :, 2]))


def test_ffn_recompute_fwd_bwd():
    """transformer MLP node that rebuilds its hidden activation in backward (ops._FFN) against fp64 autograd."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 128, 256, generator=g)
    w1 = torch.randn(1024, 256, generator=g) / 16
    b1 = torch.randn(1024, generator=g) * 0.1
    w2 = torch.randn(256, 1024, generator=g) / 32
    b2 = torch.randn(256, generator=g) * 0.1
    go = torch.randn(3, 128, 256, generator=g)
    ref_in = [t.double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    h = torch.relu(torch.nn.functional.linear(ref_in[0], ref_in[1], ref_in[2]))
    pre = torch.nn.functional.linear(h, ref_in[3], ref_in[4])
    go = go * (pre.detach().abs() > 1e-4).float()                   # stay away from the outer ReLU boundary
    torch.relu(pre).backward(go.double())
    dev_in = [t.to(DEV).requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    xd = dev_in[0]
    xs = ops.attach_split(xd, *ops.split_bf16(xd.detach()))
    y = ops.ffn(xs, *dev_in[1:], 3)
    assert y.grad_fn is not None and type(y.grad_fn).__name__.startswith('_FFN')
    y.backward(go.to(DEV))
    assert (y.detach().double().cpu() - torch.relu(pre).detach()).abs().max().item() <= 3e-5 * pre.abs().max().item()
    for got, want, n in zip(dev_in, ref_in, ['dx', 'dw1', 'db1', 'dw2', 'db2']):
        err = (got.grad.double().cpu() - want.grad).abs().max().item()
        assert err <= 2e-4 * want.grad.abs().max().item(), (n, err, want.grad.abs().max().item())


def test_linear_skinny_split_k():
    """few rows x very long reduction (the spatial encoder's fc) goes through split-K + TMA reduce-add"""
    g = torch.Generator().manual_seed(9)
    x = torch.randn(264, 8192, generator=g)
    w = torch.randn(256, 8192, generator=g) / 90
    b = torch.randn(256, generator=g)
    go = torch.randn(264, 256, generator=g)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    pre = torch.nn.functional.linear(xr, wr, br)
    go = go * (pre.detach().abs() > 1e-4).float()              # keep the gradient check away from the ReLU boundary
    ref = torch.relu(pre)
    ref.backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    y = ops.linear(xd, wd, bd, relu=True)
    y.backward(go.to(DEV))
    _check(y.detach(), ref.detach(), tol=3e-5)
    for got, want, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        err = (got.double().cpu() - want).abs().max().item()
        assert err <= 1e-3 * want.abs().max().item(), (n, err)


@pytest.mark.parametrize('mc', [1, 4])
def test_gemm_ex_exact_operands(mc):
    """a_exact / b_exact: an operand that is exactly representable in bf16 carries no lo half (two products, not three)"""
    g = torch.Generator().manual_seed(12)
    M, N, K = 2048, 256, 512
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).float()          # exact in bf16
    w = torch.randn(N, K, generator=g) / K ** 0.5
    a_hi, a_lo = _split(a)
    assert not bool(a_lo.float().any())
    w_hi, w_lo = _split(w)
    c = torch.empty(M, N, device=DEV)
    _lib.gemm_ex(a_hi=a_hi, a_lo=None, b_hi=w_hi, b_lo=w_lo, alpha=1.0, terms=3, c=c, m=M, n=N, k=K, batch=1, inner=1, splits=1,
                 a_exact=1, bn=128, mc=mc)
    _check(c, a.double() @ w.double().t())
    # exact B, MN-major operands, split-K accumulate (the weight-gradient form)
    dy = torch.randn(M, N, generator=g)
    d_hi, d_lo = _split(dy)
    gw = torch.zeros(N, K, device=DEV)
    _lib.gemm_ex(a_hi=d_hi, a_lo=d_lo, b_hi=a_hi, b_lo=None, a_mn=1, b_mn=1, alpha=1.0, terms=3, c=gw, m=N, n=K, k=M, batch=1,
                 inner=1, splits=4, c_accumulate=1, b_exact=1, bn=128, mc=mc)
    _check(gw, dy.double().t() @ a.double(), tol=5e-5)


def test_linear_64_wide_output():
    """N = 64 (the stacked key projections of the two pointer heads): 64-wide tiles forward, dX and dW on the tensor cores"""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(5, 512, 256, generator=g)
    w = torch.randn(64, 256, generator=g) / 16
    b = torch.randn(64, generator=g)
    go = torch.randn(5, 512, 64, generator=g)
    xr, wr, br = [t.double().requires_grad_(True) for t in (x, w, b)]
    torch.nn.functional.linear(xr, wr, br).backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    y = ops.linear(xd, wd, bd, relu=False, allow_n64=True)
    assert type(y.grad_fn).__name__.startswith('_SplitLinear')
    y.backward(go.to(DEV))
    _check(y.detach(), torch.nn.functional.linear(xr, wr, br).detach(), tol=3e-5)
    for got, want, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        err = (got.double().cpu() - want).abs().max().item()
        assert err <= 1e-4 * want.abs().max().item(), (n, err)


@pytest.mark.parametrize('b_mn', [0, 1])
def test_gemm_auto_pair_matches_single_cta(b_mn):
    """M large enough for the automatic CTA-pair choice (>= 74 tile pairs): bit-identical to the single-CTA kernel, including
    the ReLU-mask + column-sum epilogue of the FFN dH product (same k order, same term order, same epilogue arithmetic)."""
    M, N, K = 74 * 256 + 100, 512, 256
    g = torch.Generator().manual_seed(7 + b_mn)
    a = torch.randn(M, K, generator=g)
    w = (torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K, generator=g)) / K ** 0.5
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    mask = torch.relu(torch.randn(M, N, generator=g)).to(DEV).bfloat16()
    outs = []
    for mc in (1, 0):
        c_hi = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        c_lo = torch.empty_like(c_hi)
        colsum = torch.zeros(N, device=DEV)
        _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, alpha=1.0, terms=3, c=None, c_hi=c_hi, c_lo=c_lo, m=M, n=N, k=K,
                     batch=1, inner=1, splits=1, b_mn=b_mn, relu_mask=mask, colsum=colsum, bn=256, mc=mc)
        c = torch.empty(M, N, device=DEV)
        _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, alpha=1.0, terms=3, c=c, m=M, n=N, k=K, batch=1, inner=1,
                     splits=1, b_mn=b_mn, bn=256, mc=mc)
        torch.cuda.synchronize()
        outs.append((c_hi, c_lo, colsum, c))
    (h1, l1, s1, c1), (h0, l0, s0, c0) = outs
    assert torch.equal(c1, c0) and torch.equal(h1, h0) and torch.equal(l1, l0)
    ref = a.double() @ (w.double() if b_mn else w.double().t())
    _check(c0, ref)
    masked = ref * (mask.cpu().double() > 0)
    assert (h0.double().cpu() + l0.double().cpu() - masked).abs().max() <= 2e-5 * ref.abs().max()
    assert (s0.double().cpu() - masked.sum(0)).abs().max() <= 1e-4 * masked.abs().sum(0).max()
    assert (s1 - s0).abs().max() <= 1e-4 * s0.abs().max()          # atomics: order differs
