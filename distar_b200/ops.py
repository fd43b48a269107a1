"""Operators of the hot path: autograd wrappers around the C-ABI kernels of libdistar_b200.so.

Every operator here launches hand-written sm_100a kernels when its inputs live on a CUDA device and RAISES
otherwise — there is no silent fallback.  The only exception is the explicit host-logic test switch
``enable_host_logic_testing()`` used by the ``-m "not gpu"`` tests: it swaps each kernel for a few lines of plain
torch so the *Python* control flow around the kernels (shapes, masks, autograd wiring) can be exercised on a
machine without a GPU.  Product code never turns it on.
"""
import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import lib

_HOST_LOGIC_TESTING = False


def enable_host_logic_testing(flag: bool = True):
    """TESTS ONLY: run the torch stand-ins below instead of the CUDA kernels (CPU tensors)."""
    global _HOST_LOGIC_TESTING
    _HOST_LOGIC_TESTING = flag


def _use_kernel(t: torch.Tensor) -> bool:
    if t.is_cuda:
        return True
    if _HOST_LOGIC_TESTING:
        return False
    raise RuntimeError('distar_b200 operators need CUDA tensors (libdistar_b200.so kernels); there is no CPU path')


# ------------------------------------------------------------------------------------------------
# scatter_connection (K6)
# ------------------------------------------------------------------------------------------------
class _ScatterConnection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, project, ex, ey, entity_num, H, W):
        N, E, C = project.shape
        assert C == 32 and ex.dtype == torch.uint8 and ey.dtype == torch.uint8
        project = project.contiguous()
        out = torch.empty((N, C, H, W), dtype=torch.float32, device=project.device)
        lib.call('dsb_scatter_connection_fwd', project, ex, ey, entity_num, out, N, E, H, W)
        ctx.save_for_backward(ex, ey, entity_num)
        ctx.shape = (N, E, H, W)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ex, ey, entity_num = ctx.saved_tensors
        N, E, H, W = ctx.shape
        grad = torch.empty((N, E, 32), dtype=torch.float32, device=grad_out.device)
        lib.call('dsb_scatter_connection_bwd', grad_out.contiguous(), ex, ey, entity_num, grad, N, E, H, W)
        return grad, None, None, None, None, None


def scatter_connection(project: torch.Tensor, ex: torch.Tensor, ey: torch.Tensor, entity_num: torch.Tensor,
                       H: int, W: int) -> torch.Tensor:
    """map[n,c,y,x] = sum over valid entities at (x,y) of project[n,e,c]  -> [N,32,H,W] (contiguous NCHW).

    Reference: ``scatter_connection(..., 'add')`` model/module_utils.py:11-34 with the entity mask of
    model/encoder.py:37-38 folded in (entities >= entity_num contribute nothing).
    """
    ex, ey = ex.contiguous(), ey.contiguous()
    entity_num = entity_num.to(torch.int64).contiguous()
    if _use_kernel(project):
        return _ScatterConnection.apply(project, ex, ey, entity_num, H, W)
    N, E, C = project.shape
    valid = (torch.arange(E, device=project.device).unsqueeze(0) < entity_num.unsqueeze(1)).unsqueeze(-1)
    idx = ey.long().clamp(0, H - 1) * W + ex.long().clamp(0, W - 1)
    out = torch.zeros(N, H * W, C, dtype=project.dtype, device=project.device)
    out = out.scatter_add(1, idx.unsqueeze(-1).expand(-1, -1, C), project * valid)
    return out.view(N, H, W, C).permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------------
# return scans (K18)
# ------------------------------------------------------------------------------------------------
def return_scan(reward: torch.Tensor, value: torch.Tensor, rho: torch.Tensor, gamma_td: torch.Tensor,
                lambda_td: float = 0.8):
    """reward [F,T,B], value [F,T+1,B], rho [R,T,B] -> (vtrace_adv [F,R,T,B], upgo_adv [R,T,B], td_ret [F,T,B]).

    No gradient flows through any output (the reference wraps all three in torch.no_grad(),
    as_rl_utils.py:14,38,237).  Field 0 must be winloss (UPGO uses it, rl_loss.py:124-126).
    """
    F_, T, B = reward.shape
    R = rho.shape[0]
    reward, value, rho = reward.detach().float().contiguous(), value.detach().float().contiguous(), rho.detach().float().contiguous()
    gamma_td = gamma_td.float().contiguous()
    if _use_kernel(reward):
        vt = torch.empty((F_, R, T, B), dtype=torch.float32, device=reward.device)
        up = torch.empty((R, T, B), dtype=torch.float32, device=reward.device)
        td = torch.empty((F_, T, B), dtype=torch.float32, device=reward.device)
        lib.call('dsb_return_scan', reward, value, rho, gamma_td, float(lambda_td), vt, up, td, F_, R, T, B)
        return vt, up, td
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


# ------------------------------------------------------------------------------------------------
# per-row categorical statistics (K17) and sampling (K15)
# ------------------------------------------------------------------------------------------------
class _CategoricalStats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, teacher, action):
        rows, C = logits.shape
        dev = logits.device
        lse = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        logp = torch.empty(rows, dtype=torch.float32, device=dev)
        ent = torch.empty_like(logp)
        kl = torch.empty_like(logp) if teacher is not None else None
        lse_t = torch.empty_like(lse) if teacher is not None else None
        lib.call('dsb_categorical_stats_fwd', logits, teacher, action, lse, logp, ent, kl, lse_t, rows, C)
        ctx.save_for_backward(logits, teacher, action, lse, ent, lse_t)
        ctx.mark_non_differentiable(lse)
        if teacher is None:
            kl = torch.zeros_like(logp)
        return logp, ent, kl, lse

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_kl, _g_lse):
        logits, teacher, action, lse, ent, lse_t = ctx.saved_tensors
        rows, C = logits.shape
        grad = torch.empty_like(logits)
        cont = lambda g: None if g is None else g.contiguous().float()
        lib.call('dsb_categorical_stats_bwd', logits, teacher, action, lse, ent, lse_t, cont(g_logp), cont(g_ent),
                 cont(g_kl) if teacher is not None else None, grad, rows, C)
        return grad, None, None


def categorical_stats(logits: torch.Tensor, action: torch.Tensor, teacher: Optional[torch.Tensor] = None):
    """Per row of ``logits[..., C]``: (log p(action), entropy, KL(teacher || target)); all differentiable wrt logits.

    One fused pass instead of Categorical(...).probs/.logits/.log_prob + the entropy and KL passes of
    rl_training/rl_loss.py:63-90 and as_rl_utils.py:52-103.
    """
    shape = logits.shape[:-1]
    C = logits.shape[-1]
    z = logits.reshape(-1, C)
    t = teacher.reshape(-1, C).float() if teacher is not None else None
    a = action.reshape(-1).to(torch.int64)
    if _use_kernel(z):
        z = z.contiguous().float()
        logp, ent, kl, _ = _CategoricalStats.apply(z, t.contiguous() if t is not None else None, a.contiguous())
    else:
        lp = torch.log_softmax(z, -1)
        logp = lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)
        ent = -(lp.exp() * lp).sum(-1)
        if t is not None:
            tl = torch.log_softmax(t, -1)
            kl = (tl.exp() * (tl - lp)).sum(-1)
        else:
            kl = torch.zeros_like(logp)
    return logp.view(shape), ent.view(shape), kl.view(shape)


def sample_categorical(logits: torch.Tensor, generator: Optional[torch.Generator] = None, rng: str = 'cuda'):
    """index = argmax(softmax(logits) / q), q ~ Exp(1): torch.multinomial(p, 1)'s n=1 algorithm (K15).

    rng='cpu' draws q with ``torch.empty(shape).exponential_()`` on the CPU default generator — exactly the stream
    the reference consumes on a CPU run — and ships it to the device (parity mode); rng='cuda' draws on device.
    Returns (index int64 [rows], logp float32 [rows]).
    """
    rows, C = logits.shape
    if rng == 'cpu':
        q = torch.empty((rows, C), dtype=torch.float32).exponential_(1, generator=generator).to(logits.device)
    else:
        q = torch.empty((rows, C), dtype=torch.float32, device=logits.device).exponential_(1, generator=generator)
    if _use_kernel(logits):
        z = logits.detach().contiguous().float()
        index = torch.empty(rows, dtype=torch.int64, device=logits.device)
        logp = torch.empty(rows, dtype=torch.float32, device=logits.device)
        lib.call('dsb_sample_categorical', z, q, index, logp, rows, C)
        return index, logp
    p = torch.softmax(logits.detach().float(), -1)
    index = (p / q).argmax(-1)
    logp = torch.log_softmax(logits.detach().float(), -1).gather(-1, index.unsqueeze(-1)).squeeze(-1)
    return index, logp


# ------------------------------------------------------------------------------------------------
# split-precision tcgen05 linear (K2/K3 workhorse)
# ------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo == x to ~2^-17 relative."""
    x = x.contiguous()
    if _use_kernel(x):
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lib.call('dsb_split_bf16', x, hi, lo, x.numel())
        return hi, lo
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


# DEBUG ONLY: DSB_DISABLE_TCGEN05=1 routes every fc_block through the library matmul so the rest of the pipeline
# can be bisected on a GPU box when the tensor-core kernel is under suspicion.  Never set in tests of record/bench.
_TCGEN05_OFF = os.environ.get('DSB_DISABLE_TCGEN05', '0') == '1'


def gemm_eligible(N: int, K: int) -> bool:
    return (not _TCGEN05_OFF) and N % 128 == 0 and K % 64 == 0


def gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu: bool, terms: int = 3, want_split: bool = False):
    """C = act(A @ W^T + bias) on the tensor cores; A [M,K], W [N,K] as bf16 (hi, lo) pairs -> fp32 [M,N]."""
    M, K = a_hi.shape
    N = w_hi.shape[0]
    assert w_hi.shape[1] == K and gemm_eligible(N, K), (M, N, K)
    if _use_kernel(a_hi):
        c = torch.empty((M, N), dtype=torch.float32, device=a_hi.device)
        c_hi = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        c_lo = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        lib.call('dsb_gemm_bf16_split', a_hi, a_lo, w_hi, w_lo, bias, c, c_hi, c_lo, M, N, K, terms, 1 if relu else 0)
        return (c, c_hi, c_lo) if want_split else c
    if terms == 3:
        c = (a_hi.float() + a_lo.float()) @ (w_hi.float() + w_lo.float()).t()
    else:
        c = a_hi.float() @ w_hi.float().t()
    if bias is not None:
        c = c + bias
    if relu:
        c = torch.relu(c)
    if want_split:
        h, l = split_bf16(c)
        return c, h, l
    return c


class _SplitLinear(torch.autograd.Function):
    """y = act(x W^T + b) with forward AND input-gradient GEMMs on tcgen05 (3-term split products).

    The weight gradient dW = dY^T X reduces over the (huge) row dimension; it needs MN-major operands / split-K
    and is still a library GEMM in this round."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, terms):
        x2 = x.reshape(-1, x.shape[-1])
        a_hi, a_lo = split_bf16(x2)
        w_hi, w_lo = split_bf16(weight)
        y = gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu, terms)
        ctx.save_for_backward(x2, weight, y if relu else None)
        ctx.relu, ctx.terms, ctx.has_bias, ctx.xshape = relu, terms, bias is not None, x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight, y = ctx.saved_tensors
        gy2 = gy.reshape(-1, gy.shape[-1])
        if ctx.relu:
            gy2 = gy2 * (y > 0)
        gy2 = gy2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            N, K = weight.shape
            if gemm_eligible(K, N):
                g_hi, g_lo = split_bf16(gy2)
                wt_hi, wt_lo = split_bf16(weight.t().contiguous())
                gx = gemm_split(g_hi, g_lo, wt_hi, wt_lo, None, False, ctx.terms)
            else:
                gx = gy2 @ weight
            gx = gx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            gw = gy2.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy2.sum(0)
        return gx, gw, gb, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           terms: int = 3) -> torch.Tensor:
    """fc_block forward (ctools/torch_utils/network/nn_module.py:231-270): tcgen05 split GEMM when the shape
    tiles (N % 128 == 0, K % 64 == 0), plain library matmul for the odd small layers."""
    N, K = weight.shape
    if gemm_eligible(N, K) and (x.is_cuda or _HOST_LOGIC_TESTING) and x.numel() // K >= 1:
        return _SplitLinear.apply(x, weight, bias, relu, terms)
    _use_kernel(x)
    y = F.linear(x, weight, bias)
    return torch.relu(y) if relu else y


# ------------------------------------------------------------------------------------------------
# bilinear x2 up-sampling (location head decoder, K14)
# ------------------------------------------------------------------------------------------------
class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        x = x.contiguous()
        out = torch.empty((N, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        lib.call('dsb_upsample_bilinear2x_fwd', x, out, N * C, H, W)
        ctx.shape = (N, C, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        N, C, H, W = ctx.shape
        gin = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
        lib.call('dsb_upsample_bilinear2x_bwd', g.contiguous(), gin, N * C, H, W)
        return gin


def upsample_bilinear2x(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(x, scale_factor=2., mode='bilinear') (align_corners=False) on [N,C,H,W] fp32."""
    if _use_kernel(x):
        return _Upsample2x.apply(x.float())
    return F.interpolate(x, scale_factor=2., mode='bilinear')


# ------------------------------------------------------------------------------------------------
# flat-arena optimiser (K20)
# ------------------------------------------------------------------------------------------------
class FlatAdam:
    """clip_grad_norm_(max_norm) + Adam(betas, eps) over one contiguous fp32 arena, two kernels per step.

    Reference: RLLearner._setup_optimizer rl_learner.py:73-80 (Adam(betas=(0, 0.99), eps=1e-5)),
    GradClip 'pytorch_norm' ctools/torch_utils/grad_clip.py:141-144, applied in rl_learner.py:125,132.
    ``grad_scale`` folds DistModule.sync_gradients' division by world size (dist_helper.py:421-431).
    """

    def __init__(self, param: torch.Tensor, grad: torch.Tensor, lr: float, betas=(0.0, 0.99), eps: float = 1e-5,
                 max_norm: Optional[float] = 1.0):
        assert param.is_contiguous() and grad.is_contiguous() and param.numel() == grad.numel()
        self.param, self.grad = param, grad
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.exp_avg = torch.zeros_like(param)
        self.exp_avg_sq = torch.zeros_like(param)
        self.t = 0
        self.norm = torch.zeros(1, dtype=torch.float32, device=param.device)
        self._partial = None

    def step(self, grad_scale: float = 1.0):
        self.t += 1
        n = self.param.numel()
        if _use_kernel(self.param):
            if self._partial is None:
                self._partial = torch.empty(lib.load().dsb_sumsq_partials(), dtype=torch.float32, device=self.param.device)
            norm = None
            if self.max_norm is not None:
                lib.call('dsb_grad_norm', self.grad, n, self._partial, self.norm)
                norm = self.norm
            lib.call('dsb_adam_step', self.param, self.grad, self.exp_avg, self.exp_avg_sq, n, norm,
                     float(self.max_norm or 0.0), float(grad_scale), float(self.lr), float(self.betas[0]),
                     float(self.betas[1]), float(self.eps), int(self.t), None, None)
            return self.norm
        g = self.grad * grad_scale
        self.norm = g.norm().reshape(1)
        if self.max_norm is not None:
            g = g * torch.clamp(self.max_norm / (self.norm + 1e-6), max=1.0)
        b1, b2 = self.betas
        self.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
        self.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        denom = self.exp_avg_sq.sqrt() / math.sqrt(bc2) + self.eps
        self.param.addcdiv_(self.exp_avg, denom, value=-self.lr / bc1)
        return self.norm
