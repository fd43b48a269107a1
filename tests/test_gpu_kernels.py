"""-m gpu: each CUDA kernel (through the C-ABI) against the CPU oracle / plain fp32 torch on the same seeded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

import alphastar_ref as O
from distar_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _scatter_case(N, E, H, W, seed, big_coords=False, ragged=True):
    g = torch.Generator().manual_seed(seed)
    proj = torch.randn(N, E, 32, generator=g)
    hi = 256 if big_coords else W
    ex = torch.randint(0, hi, (N, E), generator=g).to(torch.uint8)
    ey = torch.randint(0, hi, (N, E), generator=g).to(torch.uint8)
    # force duplicates: many entities on the same pixel
    ex[:, : E // 4] = ex[:, :1]
    ey[:, : E // 4] = ey[:, :1]
    en = torch.randint(0, E + 1, (N,), generator=g) if ragged else torch.full((N,), E)
    return proj, ex, ey, en


@pytest.mark.parametrize('N,E,H,W,big', [(3, 512, 128, 128, False), (2, 512, 128, 128, True), (1, 37, 128, 128, False),
                                        (2, 512, 152, 160, True), (5, 300, 128, 128, False)])
def test_scatter_connection_bit_exact(N, E, H, W, big):
    proj, ex, ey, en = _scatter_case(N, E, H, W, seed=N * 7 + E, big_coords=big)
    mask = (torch.arange(E).unsqueeze(0) < en.unsqueeze(1)).unsqueeze(-1)
    ref = O.scatter_connection(proj * mask, ex, ey, H, W)
    out = ops.scatter_connection(proj.to(DEV), ex.to(DEV), ey.to(DEV), en.to(DEV), H, W)
    assert out.shape == (N, 32, H, W) and out.is_contiguous()
    assert torch.equal(out.cpu(), ref.contiguous()), 'scatter_connection must be bit-exact (deterministic entity order)'


def test_scatter_connection_empty_and_full():
    proj, ex, ey, _ = _scatter_case(2, 512, 128, 128, seed=1, ragged=False)
    en = torch.tensor([0, 512])
    out = ops.scatter_connection(proj.to(DEV), ex.to(DEV), ey.to(DEV), en.to(DEV), 128, 128).cpu()
    assert out[0].abs().sum() == 0
    assert torch.equal(out[1], O.scatter_connection(proj[1:], ex[1:], ey[1:], 128, 128)[0].contiguous())
    # linearity: scatter(a + b) == scatter(a) + scatter(b) up to fp32 rounding; checksum: total mass preserved
    tot = out[1].double().sum().item()
    assert abs(tot - proj[1].double().sum().item()) < 1e-2


def test_scatter_connection_backward():
    proj, ex, ey, en = _scatter_case(3, 512, 128, 128, seed=5)
    p = proj.to(DEV).requires_grad_(True)
    out = ops.scatter_connection(p, ex.to(DEV), ey.to(DEV), en.to(DEV), 128, 128)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    out.backward(g.to(DEV))
    pr = proj.clone().requires_grad_(True)
    mask = (torch.arange(512).unsqueeze(0) < en.unsqueeze(1)).unsqueeze(-1)
    O.scatter_connection(pr * mask, ex, ey, 128, 128).backward(g)
    assert torch.equal(p.grad.cpu(), pr.grad)


@pytest.mark.parametrize('F_,T,B', [(1, 32, 128), (3, 7, 5), (6, 64, 33)])
def test_return_scan_matches_oracle(F_, T, B):
    g = torch.Generator().manual_seed(T)
    reward = torch.randn(F_, T, B, generator=g) * 0.3
    reward[0] = 0
    reward[0, -1] = torch.randint(-1, 2, (B,), generator=g).float()
    value = torch.randn(F_, T + 1, B, generator=g)
    rho = torch.rand(6, T, B, generator=g).clamp(max=1) * 1.0
    rho[:, :, ::3] = 1.0
    gam = torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0, 0.997][:F_])
    vt, up, td = ops.return_scan(reward.to(DEV), value.to(DEV), rho.to(DEV), gam.to(DEV), 0.8)
    for f in range(F_):
        for h in range(6):
            ref = O.vtrace_advantages(rho[h], reward[f], value[f])
            assert torch.allclose(vt[f, h].cpu(), ref, rtol=1e-6, atol=1e-6), (f, h)
        ref = O.lambda_returns(reward[f], value[f], float(gam[f]), 0.8)
        assert torch.allclose(td[f].cpu(), ref, rtol=1e-6, atol=1e-6)
    ret = O.upgo_returns(reward[0], value[0])
    for h in range(6):
        assert torch.allclose(up[h].cpu(), rho[h] * (ret - value[0][:-1]), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('rows,C,masked', [(300, 327, False), (64, 2, False), (200, 513, True), (40, 16384, False),
                                           (33, 512, True), (5, 128, False)])
def test_categorical_stats_fwd_bwd(rows, C, masked):
    g = torch.Generator().manual_seed(rows + C)
    z = torch.randn(rows, C, generator=g) * 2
    t = torch.randn(rows, C, generator=g)
    if masked:
        k = torch.randint(1, C, (rows, 1), generator=g)
        m = torch.arange(C).unsqueeze(0) >= k
        z = z.masked_fill(m, -1e9)
        t = t.masked_fill(m, -1e9)
        z[0] = -1e9          # a fully padded selected-units step
        t[0] = -1e9
    a = torch.randint(0, C, (rows,), generator=g)
    if masked:
        a = torch.minimum(a, k.squeeze(1) - 1)
    zr = z.clone().requires_grad_(True)
    lp = torch.log_softmax(zr, -1)
    tl = torch.log_softmax(t, -1)
    r_logp = lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)
    r_ent = -(lp.exp() * lp).sum(-1)
    r_kl = (tl.exp() * (tl - lp)).sum(-1)
    w = torch.randn(3, rows, generator=g)
    (r_logp * w[0] + r_ent * w[1] + r_kl * w[2]).sum().backward()
    zd = z.to(DEV).requires_grad_(True)
    logp, ent, kl = ops.categorical_stats(zd, a.to(DEV), t.to(DEV))
    wd = w.to(DEV)
    (logp * wd[0] + ent * wd[1] + kl * wd[2]).sum().backward()
    assert torch.allclose(logp.cpu(), r_logp.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(ent.cpu(), r_ent.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(kl.cpu(), r_kl.detach(), rtol=1e-4, atol=2e-5)
    assert torch.allclose(zd.grad.cpu(), zr.grad, rtol=1e-4, atol=5e-5)   # __expf vs the host libm of whatever CPU runs the reference


@pytest.mark.parametrize('rows,C', [(64, 327), (64, 2), (32, 513), (16, 16384), (128, 128), (7, 512)])
def test_sampling_matches_torch_multinomial_stream(rows, C):
    g = torch.Generator().manual_seed(C)
    z = torch.randn(rows, C, generator=g) * 3
    if C in (513, 512):
        z = z.masked_fill(torch.arange(C).unsqueeze(0) >= torch.randint(1, C, (rows, 1), generator=g), -1e9)
    torch.manual_seed(1234)
    ref = torch.multinomial(torch.softmax(z, -1), 1)[:, 0]
    torch.manual_seed(1234)
    idx, logp = ops.sample_categorical(z.to(DEV), rng='cpu')
    assert torch.equal(idx.cpu(), ref)
    assert torch.allclose(logp.cpu(), torch.log_softmax(z, -1).gather(-1, ref.unsqueeze(-1)).squeeze(-1), atol=1e-5)


def test_split_bf16_reconstructs():
    x = torch.randn(1000003, generator=torch.Generator().manual_seed(0)) * 37
    hi, lo = ops.split_bf16(x.to(DEV))
    rec = hi.float() + lo.float()
    assert (rec.cpu() - x).abs().max() <= x.abs().max() * 2 ** -15
    assert ((rec.cpu() - x).abs() / x.abs().clamp(min=1e-6)).median() < 2 ** -16


def test_flat_adam_matches_torch():
    g = torch.Generator().manual_seed(3)
    n = 1_000_003
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.0, 0.99), eps=1e-5)
    p = p0.clone().to(DEV)
    grad = torch.zeros(n, device=DEV)
    mine = ops.FlatAdam(p, grad, lr=1e-3, betas=(0.0, 0.99), eps=1e-5, max_norm=1.0)
    for it in range(3):
        gr = torch.randn(n, generator=g) * (10.0 if it == 0 else 1e-4)
        ref.grad = gr.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        grad.copy_(gr.to(DEV))
        norm = mine.step()
        assert abs(norm.item() - norm_ref.item()) <= 1e-4 * norm_ref.item()
        assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('N,C,H,W', [(3, 128, 16, 16), (2, 64, 32, 32), (2, 32, 64, 64), (1, 5, 3, 7), (2, 4, 1, 1)])
def test_upsample_bilinear2x_fwd_bwd(N, C, H, W):
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(N, C, H, W, generator=g)
    go = torch.randn(N, C, 2 * H, 2 * W, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.interpolate(xr, scale_factor=2., mode='bilinear')
    ref.backward(go)
    xd = x.to(DEV).requires_grad_(True)
    out = ops.upsample_bilinear2x(xd)
    out.backward(go.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-6, atol=1e-6)
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('N,H,W,C', [(3, 16, 16, 128), (2, 32, 32, 64), (1, 64, 64, 32), (2, 3, 5, 7)])
def test_upsample_bilinear2x_nhwc_fwd_bwd(N, H, W, C):
    g = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(N, H, W, C, generator=g)
    go = torch.randn(N, 2 * H, 2 * W, C, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.interpolate(xr.permute(0, 3, 1, 2), scale_factor=2., mode='bilinear').permute(0, 2, 3, 1)
    ref.backward(go)
    xd = x.to(DEV).requires_grad_(True)
    out = ops.upsample_bilinear2x_nhwc(xd)
    out.backward(go.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-6, atol=1e-6)
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('rows,D,res', [(1000, 256, True), (37, 256, False), (4224, 1536, False), (300, 384, True), (64, 128, True)])
def test_layernorm_fwd_bwd(rows, D, res):
    g = torch.Generator().manual_seed(rows + D)
    x = torch.randn(rows, D, generator=g) * 3 + 1
    r = torch.randn(rows, D, generator=g) if res else None
    w = torch.randn(D, generator=g)
    b = torch.randn(D, generator=g)
    go = torch.randn(rows, D, generator=g)
    xr, wr, br = [t.double().clone().requires_grad_(True) for t in (x, w, b)]
    rr = r.double().clone().requires_grad_(True) if res else None
    ref = torch.nn.functional.layer_norm(xr + rr if res else xr, (D,), wr, br, 1e-5)
    ref.backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    rd = r.to(DEV).requires_grad_(True) if res else None
    y = ops.layer_norm(xd, wd, bd, rd, want_split=True)
    hi, lo = y._dsb_split
    y.backward(go.to(DEV))
    assert (y.double().cpu() - ref.detach()).abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert (hi.float() + lo.float() - y).abs().max().item() <= 1e-4 * ref.abs().max().item()
    for got, want, n in [(xd.grad, xr.grad, 'dx'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')] + \
            ([(rd.grad, rr.grad, 'dres')] if res else []):
        assert (got.double().cpu() - want).abs().max().item() <= 2e-5 * want.abs().max().item(), n


@pytest.mark.parametrize('B,H', [(128, 384), (5, 384), (64, 128)])
def test_lstm_cell_fwd_bwd(B, H):
    g = torch.Generator().manual_seed(B + H)
    ts = [torch.randn(B, 4 * H, generator=g), torch.randn(B, 4 * H, generator=g) * 5, torch.randn(B, H, generator=g),
          torch.randn(4 * H, generator=g), torch.randn(4 * H, generator=g), torch.randn(H, generator=g), torch.randn(H, generator=g)]
    gh, gc = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    refs = [t.double().clone().requires_grad_(True) for t in ts]
    ig, hg, c, wh, bh, wc, bc = refs
    F = torch.nn.functional
    gates = ig + F.layer_norm(hg, (4 * H,), wh, bh, 1e-5)
    i, f, gg, o = gates.chunk(4, 1)
    c2 = F.layer_norm(torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg), (H,), wc, bc, 1e-5)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    (h2 * gh.double()).sum().backward(retain_graph=True)
    (c2 * gc.double()).sum().backward()
    devs = [t.to(DEV).requires_grad_(True) for t in ts]
    hd, cd = ops.lstm_cell(*devs)
    ((hd * gh.to(DEV)).sum() + (cd * gc.to(DEV)).sum()).backward()
    assert (hd.double().cpu() - h2.detach()).abs().max().item() <= 2e-5
    assert (cd.double().cpu() - c2.detach()).abs().max().item() <= 2e-5 * c2.abs().max().item()
    for got, want in zip(devs, refs):
        assert (got.grad.double().cpu() - want.grad).abs().max().item() <= 1e-4 * want.grad.abs().max().item()


@pytest.mark.parametrize('N,H,W,C,Cpad', [(2, 64, 64, 32, 64), (3, 8, 8, 32, 32), (1, 5, 7, 3, 8)])
def test_upsample_conv3x3_single_fwd_bwd(N, H, W, C, Cpad):
    g = torch.Generator().manual_seed(H + W + C)
    x = torch.zeros(N, H, W, Cpad)
    x[..., :C] = torch.randn(N, H, W, C, generator=g)
    w = torch.randn(1, C, 3, 3, generator=g) / 3
    b = torch.randn(1, generator=g)
    go = torch.randn(N, 2 * H, 2 * W, generator=g)
    F = torch.nn.functional
    xr, wr, br = [t.double().clone().requires_grad_(True) for t in (x, w, b)]
    up = F.interpolate(xr[..., :C].permute(0, 3, 1, 2), scale_factor=2., mode='bilinear')
    ref = F.conv2d(up, wr, br, padding=1)[:, 0]
    ref.backward(go.double())
    xd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, w, b)]
    out = ops.upsample_conv3x3_single(xd, wd, bd)
    out.backward(go.to(DEV))
    assert (out.double().cpu() - ref.detach()).abs().max().item() <= 1e-5 * ref.abs().max().item()
    for got, want, n in [(xd.grad[..., :C], xr.grad[..., :C], 'dx'), (wd.grad, wr.grad, 'dw')]:
        assert (got.double().cpu() - want).abs().max().item() <= 1e-4 * want.abs().max().item(), n
    assert abs(bd.grad.item() - br.grad.item()) <= 1e-5 * go.abs().sum().item()


@pytest.mark.parametrize('N,ragged', [(3, True), (2, False)])
def test_spatial_stem_fwd_bwd(N, ragged):
    from distar_b200.synth import synth_obs
    g = torch.Generator().manual_seed(N)
    en = torch.tensor([512, 77, 300][:N]) if ragged else None
    obs = synth_obs(N, seed=5 + N, entity_num=en, hidden=False)
    sp, ent, num = obs['spatial_info'], obs['entity_info'], obs['entity_num']
    proj = torch.relu(torch.randn(N, 512, 32, generator=g))
    proj = proj * (torch.arange(512).unsqueeze(0) < num.unsqueeze(1)).unsqueeze(-1)
    w = torch.randn(32, 56, 1, 1, generator=g) / 4
    b = torch.randn(32, generator=g) / 4
    go = torch.randn(N, 64, 64, 64, generator=g)
    ops.enable_host_logic_testing(True)
    try:
        pr, wr, br = [t.clone().requires_grad_(True) for t in (proj, w, b)]
        ref = ops.spatial_stem(sp, pr, ent['x'], ent['y'], num, wr, br, 64)
        (ref * go).sum().backward()
    finally:
        ops.enable_host_logic_testing(False)
    to = lambda t: t.to(DEV)
    pd, wd, bd = [t.to(DEV).requires_grad_(True) for t in (proj, w, b)]
    out = ops.spatial_stem({k: to(v) for k, v in sp.items()}, pd, to(ent['x']), to(ent['y']), to(num), wd, bd, 64)
    hi, lo = out._dsb_split
    (out * go.to(DEV)).sum().backward()
    scale = ref.abs().max().item()
    assert (out.cpu() - ref.detach()).abs().max().item() <= 2e-5 * scale
    assert out[..., 32:].abs().max().item() == 0
    assert (hi.float() + lo.float() - out).abs().max().item() <= 1e-4 * scale
    for got, want, n in [(pd.grad, pr.grad, 'dproject'), (wd.grad, wr.grad, 'dw'), (bd.grad, br.grad, 'db')]:
        err = (got.cpu() - want).abs().max().item()
        assert err <= 2e-4 * want.abs().max().item(), (n, err, want.abs().max().item())


def test_entity_features_kernel_matches_reference_expansion():
    from distar_b200.policy_net import ENTITY_FIELDS, Net
    from distar_b200.synth import synth_obs
    obs = synth_obs(3, seed=9, hidden=False)
    ent = {k: v.clone() for k, v in obs['entity_info'].items()}
    ent['unit_type'][0, :5] = 300          # >= vocab: clamped like the reference
    ent['x'][1, :4] = 255
    ref = Net({}, 128, 128).entity_features(ent)                         # torch expansion on the CPU (pure function)
    hi, lo = ops.entity_features_split({k: v.to(DEV) for k, v in ent.items()}, ENTITY_FIELDS)
    got = hi.float() + lo.float()
    assert got.shape == (3, 512, 1024)
    assert (got.cpu() - ref).abs().max().item() <= 1e-6
    assert torch.equal(hi[..., 997:].cpu().float(), torch.zeros(3, 512, 27))
    # exact-operand layout: no lo tensor, scalar-field residuals in the spare columns; x . w is unchanged when the weight
    # repeats those fields' columns there
    xh, xl = ops.entity_features_split({k: v.to(DEV) for k, v in ent.items()}, ENTITY_FIELDS, exact=True)
    assert xl is None and torch.equal(xh[..., :997], hi[..., :997])
    ucols = [off for off, (n_, kind, wd) in zip(__import__('itertools').accumulate([0] + [f[2] for f in ENTITY_FIELDS[:-1]]),
                                                 ENTITY_FIELDS) if kind == 'u']
    assert torch.equal(xh[..., 997:997 + len(ucols)], lo[..., ucols]) and not bool(xh[..., 997 + len(ucols):].any())
    w = torch.randn(256, 997, generator=torch.Generator().manual_seed(1)).to(DEV)
    wx = ops.entity_exact_weight(w, ENTITY_FIELDS)
    want = ref.double() [..., :997] @ w.double().cpu().t()
    got = xh.double().cpu() @ wx.double().cpu().t()
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    ent['last_selected_units'][0, 0] = -1
    with pytest.raises(RuntimeError):
        ops.entity_features_split({k: v.to(DEV) for k, v in ent.items()}, ENTITY_FIELDS)


@pytest.mark.parametrize('rows,N,bf16_mask', [(1000, 64, False), (4097, 128, True), (300, 1024, True), (77, 320, False),
                                              (100000, 64, True)])
def test_relu_bwd_split_layouts(rows, N, bf16_mask):
    """column-thread x row-group layouts of the ReLU-backward pass (narrow conv activations up to the 1024-wide MLP)."""
    g = torch.Generator().manual_seed(rows + N)
    gy = torch.randn(rows, N, generator=g)
    y = torch.relu(torch.randn(rows, N, generator=g))
    want = gy * (y > 0)
    yd = y.to(DEV)
    mask = yd.to(torch.bfloat16) if bf16_mask else yd
    if bf16_mask:                                  # bf16 rounding keeps sign and zero; tiny positives may round to > 0 only
        want = gy * (mask.float().cpu() > 0)
    hi, lo, gb, gout = ops.relu_bwd_split(gy.to(DEV), mask, True, need_g=True)
    assert torch.equal(gout.cpu(), want)
    eh, el = ops.split_bf16(want.to(DEV))
    assert torch.equal(hi, eh) and torch.equal(lo, el)
    ref = want.double().sum(0)
    assert (gb.double().cpu() - ref).abs().max().item() <= 1e-5 * max(1.0, want.abs().sum(0).max().item())
    _, _, gb2, g2 = ops.relu_bwd_split(gy.to(DEV), mask, True, need_g=True, need_split=False)
    assert torch.equal(g2.cpu(), want) and torch.allclose(gb2, gb)


@pytest.mark.parametrize('with_skip', [True, False])
def test_gate_update_fwd_bwd(with_skip):
    """GatedResBlock tail (module_utils.py:228-229) fused with the next block's skip add, against torch autograd."""
    g_ = torch.Generator().manual_seed(21)
    shape = (3, 16, 16, 128)
    r, g, x, sk, go = [torch.randn(shape, generator=g_) for _ in range(5)]
    sp = torch.tensor([0.1])
    ref_in = [t.double().requires_grad_(True) for t in (r, g, x, sp, sk)]
    y = torch.tanh(ref_in[0] * torch.sigmoid(ref_in[1])) * ref_in[3] + ref_in[2]
    go = go * (y.detach().abs() > 1e-4).float()                      # stay off the ReLU boundary
    out = torch.relu(y) + (ref_in[4] if with_skip else 0)
    out.backward(go.double())
    dev_in = [t.to(DEV).requires_grad_(True) for t in (r, g, x, sp, sk)]
    got = ops.gate_update(dev_in[0], dev_in[1], dev_in[2], dev_in[3], dev_in[4] if with_skip else None)
    got.backward(go.to(DEV))
    assert (got.detach().double().cpu() - out.detach()).abs().max().item() <= 1e-5
    hi, lo = got._dsb_split
    assert torch.equal(hi, ops.split_bf16(got.detach().clone())[0]) and torch.equal(lo, ops.split_bf16(got.detach().clone())[1])
    for a, b, n in zip(dev_in, ref_in, ['dr', 'dg', 'dx', 'dsp', 'dskip']):
        if n == 'dskip' and not with_skip:
            continue
        err = (a.grad.double().cpu() - b.grad).abs().max().item()
        assert err <= 1e-4 * max(b.grad.abs().max().item(), 1e-6), (n, err)


@pytest.mark.parametrize('N,H,W,C', [(3, 64, 64, 64), (2, 32, 32, 128), (1, 2, 4, 8)])
def test_max_pool2_nhwc_fwd_bwd(N, H, W, C):
    g = torch.Generator().manual_seed(H + C)
    x = torch.relu(torch.randn(N, H, W, C, generator=g))                  # post-ReLU input, like the spatial encoder's
    go = torch.randn(N, H // 2, W // 2, C, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    ref.backward(go)
    xd = x.to(DEV).requires_grad_(True)
    out = ops.max_pool2_nhwc(xd)
    out.backward(go.to(DEV))
    assert torch.equal(out.detach().cpu(), ref.detach())
    hi, lo = out._dsb_split
    eh, el = ops.split_bf16(out.detach().clone())
    assert torch.equal(hi, eh) and torch.equal(lo, el)
    # ties only happen at 0 (both pick some zero of the window); compare the gradient where the input is positive, and
    # check that every window passes its whole gradient to exactly one position
    pos = x > 0
    assert torch.equal(xd.grad.cpu()[pos], xr.grad[pos])
    win = xd.grad.cpu().view(N, H // 2, 2, W // 2, 2, C).sum(dim=(2, 4))
    assert torch.allclose(win, go)


@pytest.mark.parametrize('L,B,H', [(5, 64, 384), (1, 3, 384), (4, 128, 128), (33, 128, 384), (3, 5, 384)])
def test_lstm_layer_matches_cell_loop(L, B, H):
    """the layer-level autograd node (one dW_hh GEMM, in-place LayerNorm gradients) against the per-cell formulation"""
    g = torch.Generator().manual_seed(L + B + H)
    G4 = 4 * H
    ig = torch.randn(L, B, G4, generator=g)
    h0, c0 = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    w = torch.randn(G4, H, generator=g) / H ** 0.5
    gh, bh = 1 + 0.1 * torch.randn(G4, generator=g), 0.1 * torch.randn(G4, generator=g)
    gc, bc = 1 + 0.1 * torch.randn(H, generator=g), 0.1 * torch.randn(H, generator=g)
    go, gcl = torch.randn(L, B, H, generator=g), torch.randn(B, H, generator=g)

    def run(layer):
        leaves = [t.to(DEV).requires_grad_(True) for t in (ig, h0, c0, w, gh, bh, gc, bc)]
        a_ig, a_h, a_c, a_w, a_gh, a_bh, a_gc, a_bc = leaves
        if layer:
            hs, cl = ops.lstm_layer(a_ig, a_h, a_c, a_w, a_gh, a_bh, a_gc, a_bc)
        else:
            h, c, ys = a_h, a_c, []
            for t in range(L):
                h, c = ops.lstm_cell(a_ig[t], h @ a_w.t(), c, a_gh, a_bh, a_gc, a_bc)
                ys.append(h)
            hs, cl = torch.stack(ys), c
        ((hs * go.to(DEV)).sum() + (cl * gcl.to(DEV)).sum()).backward()
        return hs.detach(), cl.detach(), [t.grad for t in leaves]
    hs_a, cl_a, ga = run(True)
    hs_b, cl_b, gb = run(False)
    # the recurrent product is summed in a different order than the library GEMM of the cell loop: fp32 reassociation,
    # amplified a little by every LayerNorm of the recurrence
    eh, ec = (hs_a - hs_b).abs().max().item(), (cl_a - cl_b).abs().max().item()
    assert eh <= 2e-5 * max(1.0, hs_b.abs().max().item()) and ec <= 2e-5 * max(1.0, cl_b.abs().max().item()), (eh, ec)
    for a, b, n in zip(ga, gb, ['ig', 'h0', 'c0', 'w_hh', 'gam_h', 'bet_h', 'gam_c', 'bet_c']):
        assert (a - b).abs().max().item() <= 2e-4 * max(b.abs().max().item(), 1e-6), n


@pytest.mark.parametrize('rows,C', [(300, 327), (40, 16384), (64, 2)])
def test_categorical_label_smoothing_term_and_label_flag(rows, C):
    """mean_j log p_j (LabelSmoothingCrossEntropy, sl_loss.py:16-34) from the same pass, its gradient, and the out-of-range
    label flag (torch raises on such a label; the kernel clamps and records)."""
    g = torch.Generator().manual_seed(rows * 3 + C)
    z = torch.randn(rows, C, generator=g) * 2
    a = torch.randint(0, C, (rows,), generator=g)
    zr = z.clone().requires_grad_(True)
    lp = torch.log_softmax(zr, -1)
    want = 0.9 * (-lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)) + 0.1 * (-lp.mean(-1))
    w = torch.randn(rows, generator=g)
    (want * w).sum().backward()
    zd = z.to(DEV).requires_grad_(True)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    logp, _, _, mean_lp = ops.categorical_stats(zd, a.to(DEV), want_mean=True, flag=flag)
    got = 0.9 * (-logp) + 0.1 * (-mean_lp)
    (got * w.to(DEV)).sum().backward()
    assert torch.allclose(got.cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(zd.grad.cpu(), zr.grad, rtol=1e-4, atol=5e-5)
    assert int(flag.item()) == 0
    bad = a.clone()
    bad[rows // 2] = C
    bad[0] = -1
    lp2 = ops.categorical_stats(z.to(DEV), bad.to(DEV), flag=flag)[0]
    assert int(flag.item()) == 2 and bool(torch.isfinite(lp2).all())
    # a logit of -inf is a legal mask value for the kernel (the reference itself uses -1e9)
    zi = z.clone()
    zi[:, 0] = float('-inf')
    a1 = a.clamp(min=1)
    li = ops.categorical_stats(zi.to(DEV), a1.to(DEV))[0]
    assert torch.allclose(li.cpu(), torch.log_softmax(zi, -1).gather(-1, a1.unsqueeze(-1)).squeeze(-1), rtol=1e-5, atol=1e-5)


def test_flat_adam_weight_decay_schedule_and_skip_flag():
    g = torch.Generator().manual_seed(5)
    n = 300_001
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p = p0.clone().to(DEV)
    grad = torch.zeros(n, device=DEV)
    mine = ops.FlatAdam(p, grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.4, clip_type='momentum_norm')
    sched, sched_ref = (torch.optim.lr_scheduler.MultiStepLR(o, milestones=[2], gamma=0.1) for o in (mine, opt))
    for it in range(4):
        gr = torch.randn(n, generator=g) * 3
        ref.grad = gr.clone()
        opt.step()
        sched_ref.step()
        grad.copy_(gr.to(DEV))
        norm = mine.step(grad_scale=0.5) if it == 1 else mine.step()
        sched.step()
        if it == 1:                                   # grad_scale = 1/world: the norm reported is the averaged gradient's
            assert abs(norm.item() - 0.5 * gr.norm().item()) <= 1e-4 * gr.norm().item()
            ref.data.copy_(p.cpu())                   # (the reference side did not scale: re-synchronise the two)
            opt.state[ref]['exp_avg'].copy_(mine.exp_avg.cpu())
            opt.state[ref]['exp_avg_sq'].copy_(mine.exp_avg_sq.cpu())
        else:
            assert abs(norm.item() - gr.norm().item()) <= 1e-4 * gr.norm().item()      # 'momentum_norm' never clips
            assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-6), it
    before, m_before = p.clone(), mine.exp_avg.clone()
    grad.normal_()
    mine.step(skip_flag=torch.ones(4, device=DEV))
    assert torch.equal(p, before) and torch.equal(mine.exp_avg, m_before)


@pytest.mark.parametrize('M,K,N,relu,dtype', [(264, 10, 64, True, torch.float32), (264, 90, 128, True, torch.int16),
                                             (37, 260, 128, True, torch.uint8), (4096, 256, 327, False, torch.float32),
                                             (1000, 256, 2, False, torch.float32), (513, 32, 256, True, torch.float32),
                                             (300, 1024, 32, True, torch.float32), (77, 32, 32, False, torch.float32),
                                             (4224, 256, 1, False, torch.float32), (1, 269, 64, True, torch.uint8)])
def test_linear_any_shapes_fwd_bwd(M, K, N, relu, dtype):
    """fc_block shapes the tile grid does not divide (scalar encoder, head MLPs, value_fc) on the tcgen05 kernel through
    zero padding: forward, input gradient and weight / bias gradients against fp64."""
    g = torch.Generator().manual_seed(M + K + N)
    if dtype == torch.float32:
        x = torch.randn(M, K, generator=g)
    else:
        x = torch.randint(0, 21 if dtype == torch.uint8 else 2, (M, K), generator=g).to(dtype)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    go = torch.randn(M, N, generator=g)
    xd = x.to(DEV).requires_grad_(dtype == torch.float32)
    wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    out = ops.linear(xd, wd, bd, relu, 3, exact_input=dtype != torch.float32)
    out.backward(go.to(DEV))
    xr = x.double().requires_grad_(dtype == torch.float32)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    pre = F.linear(xr, wr, br)
    # the reference takes the ReLU decisions of the device result: a pre-activation within rounding of 0 (0/1 inputs make exact
    # cancellations likely) may legitimately fall on either side, and one flipped decision moves a gradient entry by O(1)
    ref = pre * (out.detach().cpu() > 0) if relu else pre
    assert not relu or ((pre.detach() > 0) != (out.detach().cpu() > 0)).float().mean().item() < 1e-3
    ref.backward(go.double())
    scale = ref.abs().max().item()
    assert out.shape == (M, N)
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() <= 2e-5 * max(scale, 1.0)
    gs = wr.grad.abs().max().item()
    assert (wd.grad.cpu().double() - wr.grad).abs().max().item() <= 1e-4 * max(gs, 1e-3)
    assert (bd.grad.cpu().double() - br.grad).abs().max().item() <= 1e-4 * max(br.grad.abs().max().item(), 1e-3)
    if dtype == torch.float32:
        assert (xd.grad.cpu().double() - xr.grad).abs().max().item() <= 1e-4 * max(xr.grad.abs().max().item(), 1e-3)


def test_linear_any_accumulates_into_arena_slots():
    """when weight / bias are arena leaves their gradients are ADDED in place (TMA reduce-add clipped to [N, K])."""
    g = torch.Generator().manual_seed(0)
    M, K, N = 300, 32, 200
    x = torch.randn(M, K, generator=g).to(DEV)
    w = torch.nn.Parameter(torch.randn(N, K, generator=g).to(DEV))
    b = torch.nn.Parameter(torch.randn(N, generator=g).to(DEV))
    w.grad, b.grad = torch.ones_like(w), torch.ones_like(b)
    go = torch.randn(M, N, generator=g).to(DEV)
    ops.linear(x, w, b, False, 3).backward(go)
    assert torch.allclose(w.grad, 1 + go.t() @ x, rtol=1e-4, atol=1e-3)
    assert torch.allclose(b.grad, 1 + go.sum(0), rtol=1e-4, atol=1e-3)


def test_glu_gate_fwd_bwd():
    g = torch.Generator().manual_seed(4)
    a, x, go = (torch.randn(777, 256, generator=g) for _ in range(3))
    ar, xr = a.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (torch.sigmoid(ar) * xr).backward(go)
    ad, xd = a.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    out = ops.glu_gate(ad, xd)
    out.backward(go.to(DEV))
    assert torch.allclose(out.cpu(), torch.sigmoid(a) * x, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ad.grad.cpu(), ar.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('embedding', [False, True])
def test_onehot_linear_fwd_bwd(embedding):
    g = torch.Generator().manual_seed(6)
    P, N, C = 500, 256, 327
    w = torch.randn((C, N) if embedding else (N, C), generator=g)
    b = None if embedding else torch.randn(N, generator=g)
    idx = torch.randint(0, C, (P,), generator=g)
    go = torch.randn(P, N, generator=g)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if b is not None else None
    ref = F.embedding(idx, wr) if embedding else F.linear(F.one_hot(idx, C).float(), wr, br)
    ref = torch.relu(ref)
    ref.backward(go)
    wd = w.to(DEV).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True) if b is not None else None
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.onehot_linear(wd, bd, idx.to(DEV), relu=True, embedding=embedding, flag=flag)
    out.backward(go.to(DEV))
    assert torch.equal(out.cpu(), ref.detach()) and int(flag.item()) == 0
    assert torch.allclose(wd.grad.cpu(), wr.grad, rtol=1e-4, atol=1e-5)
    if b is not None:
        assert torch.allclose(bd.grad.cpu(), br.grad, rtol=1e-4, atol=1e-4)
    # ids outside the vocabulary: clamped; recorded unless the caller asked for the scalar encoder's clamp-max semantics
    bad = idx.clone()
    bad[3] = C + 5
    o2 = ops.onehot_linear(wd, bd, bad.to(DEV), relu=True, embedding=embedding, clamp_max=True, flag=flag)
    assert int(flag.item()) == 0 and torch.equal(o2[3].cpu(), torch.relu((w[C - 1] if embedding else w[:, C - 1] + b)))
    ops.onehot_linear(wd, bd, bad.to(DEV), relu=True, embedding=embedding, flag=flag)
    assert int(flag.item()) == 4


@pytest.mark.parametrize('P,E,ld,col', [(50, 512, 64, 32), (3, 512, 32, 0), (7, 64, 64, 0)])
def test_target_unit_logits_fwd_bwd(P, E, ld, col):
    g = torch.Generator().manual_seed(P + E)
    kfull = torch.randn(P, E, ld, generator=g)
    q = torch.randn(P, 32, generator=g)
    en = torch.randint(1, E + 1, (P,), generator=g)
    en[0] = E
    T = 0.8
    go = torch.randn(P, E, generator=g)
    kr, qr = kfull.clone().requires_grad_(True), q.clone().requires_grad_(True)
    valid = torch.arange(E).unsqueeze(0) < en.unsqueeze(1)
    ref = torch.matmul(kr[..., col:col + 32], qr.unsqueeze(-1)).squeeze(-1).masked_fill(~valid, -1e9) / T
    ref.backward(go)
    kd, qd = kfull.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    out = ops.target_unit_logits(kd, col, qd, en.to(DEV), T)
    out.backward(go.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-5, atol=1e-4)
    assert torch.equal(out.cpu() < -1e8, ~valid)
    assert torch.allclose(kd.grad.cpu(), kr.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(qd.grad.cpu(), qr.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('rows,D', [(1000, 64), (37, 32), (5, 96), (84480, 64)])
def test_layernorm_small_fwd_bwd(rows, D):
    g = torch.Generator().manual_seed(rows + D)
    x, go = torch.randn(rows, D, generator=g), torch.randn(rows, D, generator=g)
    w, b = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    F.layer_norm(xr, (D,), wr, br, 1e-5).backward(go.double())
    xd, wd, bd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y = ops.layer_norm(xd, wd, bd)
    y.backward(go.to(DEV))
    assert torch.allclose(y.detach().cpu().double(), F.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-5), rtol=1e-5, atol=1e-5)
    assert torch.allclose(xd.grad.cpu().double(), xr.grad, rtol=1e-4, atol=1e-5)
    assert (wd.grad.cpu().double() - wr.grad).abs().max().item() <= 1e-4 * wr.grad.abs().max().item()
    assert (bd.grad.cpu().double() - br.grad).abs().max().item() <= 1e-4 * br.grad.abs().max().item()


@pytest.mark.parametrize('B,S,H,hd', [(50, 20, 2, 8), (3, 32, 2, 8), (7, 5, 1, 16)])
def test_small_attention_fwd_bwd(B, S, H, hd):
    g = torch.Generator().manual_seed(B + S)
    qkv, go = torch.randn(B, S, 3 * H * hd, generator=g), torch.randn(B, S, H * hd, generator=g)
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    ref = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd), dim=-1), v).permute(0, 2, 1, 3).reshape(B, S, H * hd)
    ref.backward(go.double())
    qd = qkv.to(DEV).requires_grad_(True)
    out = ops.small_attention(qd, H, hd)
    out.backward(go.to(DEV))
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(qd.grad.cpu().double(), qr.grad, rtol=1e-3, atol=1e-5)


def test_bo_tokens_match_reference_features():
    g = torch.Generator().manual_seed(2)
    B, L, W = 9, 20, 128
    bo = torch.randint(0, 174, (B, L), generator=g).to(torch.int16)
    loc = torch.randint(0, W * W, (B, L), generator=g).to(torch.int16)
    bits = torch.arange(9, -1, -1)
    want = torch.cat([F.one_hot(bo.long(), 174).float(), torch.eye(L).unsqueeze(0).expand(B, -1, -1),
                      (((loc.long() % W).unsqueeze(-1) >> bits) & 1).float(), (((loc.long() // W).unsqueeze(-1) >> bits) & 1).float()], dim=2)
    hi = ops.bo_tokens(bo.to(DEV), loc.to(DEV), W)
    assert hi.shape == (B * L, 256)
    assert torch.equal(hi.float().cpu()[:, :214], want.reshape(B * L, 214)) and float(hi.float()[:, 214:].abs().sum()) == 0


def test_pointer_training_kernels_match_torch_formulation():
    """su_prefix_mean / su_lstm / su_logits (csrc/su_train.cu) against the one-hot / cummax formulation of round 1, including
    repeated labels, rows without a selection and the shared key-gradient buffer."""
    g = torch.Generator().manual_seed(11)
    P, E, S = 7, 512, 6
    en = torch.tensor([512, 40, 300, 5, 64, 512, 9])
    num = torch.tensor([6, 3, 0, 2, 5, 4, 3])
    su = torch.zeros(P, 64, dtype=torch.long)
    for p in range(P):
        k = int(num[p])
        if k:
            units = torch.randperm(int(en[p]), generator=g)[:k - 1]
            su[p, :k - 1] = units
            su[p, k - 1] = en[p]
    su[4, 2] = su[4, 0]                                           # a repeated label
    kfull = torch.randn(P, E, 64, generator=g)
    end = torch.randn(1, 32, generator=g)
    hs = torch.randn(P, S, 32, generator=g)
    gm, gl = torch.randn(P, S, 32, generator=g), torch.randn(P, S, E + 1, generator=g)
    # ---- torch formulation (policy_net CPU path)
    kr, er, hr = kfull.clone().requires_grad_(True), end.clone().requires_grad_(True), hs.clone().requires_grad_(True)
    slot = torch.arange(E + 1).unsqueeze(0)
    key = torch.where((slot == en.unsqueeze(1)).unsqueeze(-1), er.view(1, 1, -1), F.pad(kr[..., :32], (0, 0, 0, 1)))
    valid = slot < (en + 1).unsqueeze(1)
    sus = su[:, :S]
    onehot = sus.unsqueeze(-1) == slot.unsqueeze(1)
    ended = torch.cummax((sus == en.unsqueeze(1)).long(), dim=1)[0].bool()
    picked = torch.cummax((onehot & ~ended.unsqueeze(-1)).long(), dim=1)[0].float()
    ssum = torch.matmul(picked, key)
    mean_ref = torch.where((num != 0).view(P, 1, 1), ssum / picked.sum(-1, keepdim=True), ssum)
    chosen = torch.cummax(onehot.long(), dim=1)[0].bool()
    chosen = torch.cat([torch.zeros_like(chosen[:, :1]), chosen[:, :-1]], dim=1)
    mask = valid.unsqueeze(1) & ~chosen
    mask[:, 0] = valid & (slot != en.unsqueeze(1))
    logits_ref = torch.matmul(hr, key.transpose(1, 2)).masked_fill(~mask, -1e9)
    fin = torch.isfinite(mean_ref)
    ((mean_ref * gm)[fin].sum() + (logits_ref * gl * mask).sum()).backward()
    # ---- kernels
    kd, ed, hd_ = kfull.to(DEV).requires_grad_(True), end.to(DEV).requires_grad_(True), hs.to(DEV).requires_grad_(True)
    kf, sink = ops.fork_keys(kd)
    mean = ops.su_prefix_mean(kf, su.to(DEV), en.to(DEV), num.to(DEV), S, sink)
    logits = ops.su_logits(hd_, kf, ed, su.to(DEV), en.to(DEV), sink)
    assert torch.equal(torch.isfinite(mean).cpu(), fin)
    assert torch.allclose(mean.detach().cpu()[fin], mean_ref.detach()[fin], rtol=1e-5, atol=1e-6)
    assert torch.equal(logits.detach().cpu() < -1e8, ~mask)
    assert torch.allclose(logits.detach().cpu()[mask], logits_ref.detach()[mask], rtol=1e-5, atol=1e-4)
    ((mean * gm.to(DEV))[fin.to(DEV)].sum() + (logits * gl.to(DEV) * mask.to(DEV)).sum()).backward()
    assert torch.allclose(kd.grad.cpu(), kr.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(ed.grad.cpu(), er.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(hd_.grad.cpu(), hr.grad, rtol=1e-4, atol=1e-4)
    # ---- the 32-wide LN-LSTM against a loop of the fused cell
    ig = torch.randn(P, S, 128, generator=g)
    w = torch.randn(128, 32, generator=g) / 6
    gh, bh = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    gc, bc = 1 + 0.1 * torch.randn(32, generator=g), 0.1 * torch.randn(32, generator=g)
    go = torch.randn(P, S, 32, generator=g)

    def run(kernel):
        leaves = [t.clone().double().requires_grad_(True) if not kernel else t.to(DEV).requires_grad_(True) for t in (ig, w, gh, bh, gc, bc)]
        a_ig, a_w, a_gh, a_bh, a_gc, a_bc = leaves
        if kernel:
            out = ops.su_lstm(a_ig, a_w, a_gh, a_bh, a_gc, a_bc)
            out.backward(go.to(DEV))
        else:
            h = torch.zeros(P, 32, dtype=torch.double)
            c = torch.zeros(P, 32, dtype=torch.double)
            ys = []
            for i in range(S):
                gates = a_ig[:, i] + F.layer_norm(h @ a_w.t(), (128,), a_gh, a_bh, 1e-5)
                gi, gf, gg, g_o = gates.chunk(4, 1)
                c = F.layer_norm(torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg), (32,), a_gc, a_bc, 1e-5)
                h = torch.sigmoid(g_o) * torch.tanh(c)
                ys.append(h)
            out = torch.stack(ys, dim=1)
            out.backward(go.double())
        return out.detach().cpu().double(), [t.grad.cpu().double() for t in leaves]
    oa, ga = run(True)
    ob, gb = run(False)
    assert torch.allclose(oa, ob, rtol=1e-4, atol=1e-5)
    for a, b, n in zip(ga, gb, ['ig', 'w_hh', 'gam_h', 'bet_h', 'gam_c', 'bet_c']):
        assert (a - b).abs().max().item() <= 2e-4 * max(b.abs().max().item(), 1e-6), n
