"""TEST INFRASTRUCTURE: plain-torch stand-ins for the distar_b200 operators that ALWAYS run a CUDA kernel on the GPU.

``distar_b200.ops.enable_host_logic_testing(True)`` (tests only) imports this module by name and routes CPU tensors here, so
the ``-m "not gpu"`` tests can exercise the product's HOST logic (shapes, masks, autograd wiring, learner / DP plumbing) without
a GPU, and the ``-m gpu`` kernel tests can use the same functions as their plain-torch reference.  Nothing in the package, in
bench.py or in __graft_entry__ imports it: without the test switch a CPU tensor raises in ``ops._use_kernel``."""
import torch
import torch.nn.functional as F


def scatter_connection(project, ex, ey, entity_num, H, W):
    """module_utils.py:11-34 ('add') with the entity mask of encoder.py:37-38 folded in."""
    N, E, C = project.shape
    valid = (torch.arange(E, device=project.device).unsqueeze(0) < entity_num.unsqueeze(1)).unsqueeze(-1)
    idx = ey.long().clamp(0, H - 1) * W + ex.long().clamp(0, W - 1)
    out = torch.zeros(N, H * W, C, dtype=project.dtype, device=project.device)
    out = out.scatter_add(1, idx.unsqueeze(-1).expand(-1, -1, C), project * valid)
    return out.view(N, H, W, C).permute(0, 3, 1, 2).contiguous()


def spatial_stem(planes, effects, project, ex, ey, entity_num, weight, bias, out_c):
    """spatial_encoder.py:51-79 up to the first max-pool, channels-last, channels zero-padded to out_c."""
    N, H, W = planes[0].shape
    scatter_map = scatter_connection(project, ex, ey, entity_num, H, W)
    chans = [planes[0].float().unsqueeze(1) / 256]
    for p, n in zip(planes[1:], (4, 2, 5, 2, 2, 2)):
        chans.append(F.one_hot(p.long(), n).permute(0, 3, 1, 2).float())
    for e in effects:
        m = torch.zeros(N, H * W, device=project.device)
        m.scatter_(1, e.long(), 1.0)
        chans.append(m.view(N, 1, H, W))
    chans.append(scatter_map)
    x = torch.relu(F.conv2d(torch.cat(chans, dim=1), weight, bias))
    x = F.max_pool2d(x, 2, 2).permute(0, 2, 3, 1)
    return F.pad(x, (0, out_c - 32)).contiguous()


def return_scan(reward, value, rho, gamma_td, lambda_td):
    """V-trace advantages, UPGO returns and TD(lambda) returns (as_rl_utils.py:157-312), time loops spelled out."""
    F_, T, B = reward.shape
    R = rho.shape[0]
    vt = torch.empty((F_, R, T, B))
    up = torch.empty((R, T, B))
    td = torch.empty((F_, T, B))
    for f in range(F_):
        v, r = value[f], reward[f]
        for h in range(R):
            c = rho[h]
            vs = v[T].clone()
            for t in reversed(range(T)):
                vt[f, h, t] = c[t] * (r[t] + vs - v[t])
                vs = v[t] + c[t] * (r[t] + v[t + 1] - v[t]) + c[t] * (vs - v[t + 1])
        g = gamma_td[f]
        nxt = None
        for t in reversed(range(T)):
            nxt = r[t] + g * v[t + 1] if t == T - 1 else r[t] + g * lambda_td * nxt + (g - g * lambda_td) * v[t + 1]
            td[f, t] = nxt
    v, r = value[0], reward[0]
    nxt = None
    for t in reversed(range(T)):
        if t == T - 1:
            nxt = r[t] + v[t + 1]
        else:
            lam = ((r[t + 1] + v[t + 2]) >= v[t + 1]).float()
            nxt = r[t] + lam * nxt + (1 - lam) * v[t + 1]
        up[:, t] = rho[:, t] * (nxt - v[t])
    return vt, up, td


def categorical_stats(z, t, a, flag):
    """rows of logits z [rows, C], optional teacher logits t, labels a -> (log p(a), entropy, KL(teacher || target), mean log p)."""
    C = z.shape[-1]
    if flag is not None:
        flag |= 2 * int(((a < 0) | (a >= C)).any())
    a = a.clamp(0, C - 1)
    lp = torch.log_softmax(z, -1)
    logp = lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)
    ent = -(lp.exp() * lp).sum(-1)
    mean_lp = lp.mean(-1)
    if t is not None:
        tl = torch.log_softmax(t, -1)
        kl = (tl.exp() * (tl - lp)).sum(-1)
    else:
        kl = torch.zeros_like(logp)
    return logp, ent, kl, mean_lp


def sample_categorical(logits, q):
    """torch.multinomial(softmax(logits), 1)'s n = 1 algorithm given its Exp(1) variates q."""
    p = torch.softmax(logits.detach().float(), -1)
    index = (p / q).argmax(-1)
    logp = torch.log_softmax(logits.detach().float(), -1).gather(-1, index.unsqueeze(-1)).squeeze(-1)
    return index, logp


def split_bf16(x):
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu, terms):
    if terms == 3:
        c = (a_hi.float() + (a_lo.float() if a_lo is not None else 0)) @ (w_hi.float() + w_lo.float()).t()
    else:
        c = a_hi.float() @ w_hi.float().t()
    if bias is not None:
        c = c + bias
    return torch.relu(c) if relu else c


def onehot_linear(weight, bias, i64, relu, embedding, clamp_max, flag):
    C = weight.shape[0] if embedding else weight.shape[1]
    if flag is not None:
        flag |= 4 * int(((i64 < 0) | ((i64 >= C) & (not clamp_max))).any())
    i64 = i64.clamp(0, C - 1)
    out = weight[i64] if embedding else weight.t()[i64]
    if bias is not None:
        out = out + bias
    return torch.relu(out) if relu else out


def upsample_bilinear2x_nhwc(x):
    return F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2., mode='bilinear').permute(0, 2, 3, 1).contiguous()


def upsample_bilinear2x(x):
    return F.interpolate(x, scale_factor=2., mode='bilinear')


def upshift9(z, bias):
    """sum of the nine tap-shifted x2 up-samplings of z [N,H,W,9] -> [N,2H,2W] (the last stage of the location head)."""
    up = F.interpolate(z.permute(0, 3, 1, 2), scale_factor=2., mode='bilinear')          # [N,9,2H,2W]
    up = F.pad(up, (1, 1, 1, 1))
    H2, W2 = up.shape[2] - 2, up.shape[3] - 2
    out = sum(up[:, t, t // 3:t // 3 + H2, t % 3:t % 3 + W2] for t in range(9))
    return out + bias if bias is not None else out


def flat_adam_step(opt, grad_scale, skip_flag, lr, b1, b2, eps, wd):
    """ops.FlatAdam.step on CPU tensors: clip by global norm ('pytorch_norm', grad_clip.py:141-144) + Adam with coupled weight
    decay, exactly what dsb_grad_norm + dsb_adam_step do."""
    import math
    if skip_flag is not None and float(skip_flag.reshape(-1)[0]) != 0.0:
        return opt.norm
    g = opt.grad * grad_scale
    opt.norm = g.norm().reshape(1)
    if opt.max_norm is not None:
        g = g * torch.clamp(opt.max_norm / (opt.norm + 1e-6), max=1.0)
    if wd != 0.0:
        g = g + wd * opt.param
    opt.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    opt.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** opt.t, 1 - b2 ** opt.t
    denom = opt.exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    opt.param.addcdiv_(opt.exp_avg, denom, value=-lr / bc1)
    return opt.norm


def expand_ragged(src, off, steps, width, rows, S, W, fill):
    """dst[r, s, e] = src[off[r] + s * width[r] + e] inside (steps[r], width[r]), `fill` outside (distar_b200.batch)."""
    s_idx = torch.arange(S).view(1, S, 1)
    e_idx = torch.arange(W).view(1, 1, W)
    st = steps.view(-1, 1, 1) if steps is not None else torch.ones(rows, 1, 1, dtype=torch.long)
    inside = (s_idx < st) & (e_idx < width.view(-1, 1, 1))
    idx = (off.view(-1, 1, 1) + s_idx * width.view(-1, 1, 1) + e_idx).clamp(0, max(src.numel() - 1, 0))
    gathered = src[idx] if src.numel() else torch.zeros((rows, S, W), dtype=src.dtype)
    return torch.where(inside, gathered, torch.full((), fill, dtype=src.dtype))


def sequence_mask(lengths, add, W):
    return torch.arange(W).unsqueeze(0) < (lengths + add).unsqueeze(1)


def unpack_planes(hw, packed_planes):
    v = hw.to(torch.int32) & 0xFFFF
    return {name: ((v >> bit) & ((1 << bits) - 1)).to(torch.uint8) for name, bit, bits in packed_planes}


TABLE = {f.__name__: f for f in (expand_ragged, sequence_mask, unpack_planes, upsample_bilinear2x_nhwc, upsample_bilinear2x, upshift9, flat_adam_step, scatter_connection, spatial_stem, return_scan, categorical_stats, sample_categorical, split_bf16,
                                 gemm_split, onehot_linear)}
