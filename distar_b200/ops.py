"""Operators of the hot path: autograd wrappers around the C-ABI kernels of libdistar_b200.so.

Every operator here launches hand-written sm_100a kernels when its inputs live on a CUDA device and RAISES
otherwise — there is no silent CPU fallback.  The only exception is the explicit host-logic test switch
``enable_host_logic_testing()`` used by the tests: CPU tensors then take plain-torch stand-ins so the *Python* control flow
around the kernels (shapes, masks, autograd wiring) can be exercised on a machine without a GPU.  Product code never turns
it on.  The stand-ins of the operators that always run a kernel on the GPU live in tests/host_standins.py; torch code that is
still in this file after an ``if _use_kernel(...)`` is (mostly) also the library path a CUDA tensor takes for a shape the
kernel does not cover (e.g. ``conv_geometry_supported``).
"""
import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import lib

_HOST_LOGIC_TESTING = False
_STANDINS = None          # tests only: operator name -> torch stand-in, loaded from tests/host_standins.py


def enable_host_logic_testing(flag: bool = True):
    """TESTS ONLY: CPU tensors run torch stand-ins instead of raising.  The stand-ins of the operators that ALWAYS take a kernel
    on the GPU live outside the package, in tests/host_standins.py (imported by name: the tests directory is on sys.path under
    pytest); most of what remains in this file after an `if _use_kernel(...)` is library code a CUDA tensor can also reach
    (shapes a kernel does not cover)."""
    global _HOST_LOGIC_TESTING, _STANDINS
    _HOST_LOGIC_TESTING = flag
    if flag and _STANDINS is None:
        import importlib
        _STANDINS = importlib.import_module('host_standins').TABLE


def _standin(name: str):
    return _STANDINS[name]


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
    return _standin('scatter_connection')(project, ex, ey, entity_num, H, W)


# ------------------------------------------------------------------------------------------------
# entity feature expansion (K1)
# ------------------------------------------------------------------------------------------------
_DTYPE_CODE = {torch.uint8: 0, torch.int16: 1, torch.int8: 2, torch.float16: 3}


def entity_features_split(entity_info: dict, fields, check_negative: bool = True, flag: Optional[torch.Tensor] = None,
                          exact: bool = False):
    """fields: [(name, kind 'o'|'b'|'u', width)] in concat order.  Returns the [N, E, 1024] bf16 (hi, lo) feature pair the
    embedding GEMM consumes, built straight from the wire-format fields (no fp32 concat).  None on unsupported dtypes.
    With ``flag`` (device int32[1]) the negative-id error is only recorded there and the caller raises later: reading it
    here would stall the host once per encoder chunk and drain the launch queue.
    ``exact``: one-hot / binary columns are exact in bf16 and the bf16 residuals of the scalar fields are placed in the spare
    columns [row width, row width + #scalar fields) of the hi tensor, so NO lo tensor exists (returns (hi, None)); the caller
    must repeat the scalar fields' weight columns there (``entity_exact_weight``) and run the GEMM with a_exact."""
    first = entity_info[fields[0][0]]
    if not first.is_cuda:
        return None
    tensors, kinds, offs, vocabs, dts = [], [], [], [], []
    off = 0
    for name, kind, w in fields:
        t = entity_info[name]
        if t.dtype not in _DTYPE_CODE:
            return None
        tensors.append(t.contiguous())
        kinds.append({'o': 0, 'b': 1, 'u': 2}[kind])
        offs.append(off)
        vocabs.append(w)
        dts.append(_DTYPE_CODE[t.dtype])
        off += w
    N, E = first.shape
    hi = torch.empty((N, E, 1024), dtype=torch.bfloat16, device=first.device)
    lo = None if exact else torch.empty((N, E, 1024), dtype=torch.bfloat16, device=first.device)
    deferred = flag is not None
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=first.device)
    lib.call('dsb_entity_features', lib.ptr_array(tensors), lib.int_array(kinds), lib.int_array(offs),
             lib.int_array(vocabs), lib.int_array(dts), len(fields), hi, lo, off if exact else -1, N * E, flag)
    if check_negative and not deferred and int(flag.item()) != 0:          # entity_encoder.py:69-72 raises on negative ids
        raise RuntimeError('negative categorical id in an entity field')
    return hi, lo


_CONST_IDX = {}


def entity_exact_weight(w: torch.Tensor, fields) -> torch.Tensor:
    """[out, 997] embedding weight -> [out, 1024] for the exact-operand feature rows of entity_features_split(exact=True):
    the columns of the scalar ('u') fields are repeated after the row (they multiply the bf16 residuals), zeros after."""
    off, ucols = 0, []
    for _name, kind, wd in fields:
        if kind == 'u':
            ucols.append(off)
        off += wd
    key = (w.device, tuple(ucols))
    idx = _CONST_IDX.get(key)          # built once per device: a host -> device copy here would be illegal under CUDA-graph capture
    if idx is None:
        idx = _CONST_IDX[key] = torch.tensor(ucols, device=w.device)
    return torch.cat([w, w.index_select(1, idx), w.new_zeros(w.shape[0], 1024 - off - len(ucols))], dim=1)


def linear_presplit(x_hi: torch.Tensor, x_lo: Optional[torch.Tensor], weight: torch.Tensor, bias, relu: bool, terms: int = 3,
                    emit_split: bool = False) -> torch.Tensor:
    """fc_block on an input that only exists as a bf16 (hi, lo) pair (no gradient flows to it).  x_lo None: the input is
    exactly representable in bf16 (one product less)."""
    y, y_hi, y_lo, _ = _SplitLinear.apply(x_hi, weight, bias, relu, terms, x_hi, x_lo, emit_split)
    return attach_split(y, y_hi, y_lo) if emit_split else y


# ------------------------------------------------------------------------------------------------
# fused spatial stem: scatter + plane expansion + 1x1 conv + ReLU + 2x2 max-pool (K6/K7 first stage)
# ------------------------------------------------------------------------------------------------
STEM_PLANES = ['height_map', 'visibility_map', 'creep', 'player_relative', 'alerts', 'pathable', 'buildable']
STEM_EFFECTS = ['effect_PsiStorm', 'effect_NukeDot', 'effect_LiberatorDefenderZone', 'effect_BlindingCloud',
                'effect_CorrosiveBile', 'effect_LurkerSpines']


class _SpatialStem(torch.autograd.Function):
    @staticmethod
    def forward(ctx, project, weight, bias, ex, ey, entity_num, out_c, *maps):
        planes, effects = list(maps[:7]), list(maps[7:13])
        N, H, W = planes[0].shape
        E = project.shape[1]
        dev = project.device
        project = project.contiguous()
        w2 = weight.reshape(32, 56).contiguous()
        out = torch.empty((N, H // 2, W // 2, out_c), dtype=torch.float32, device=dev)
        hi = torch.empty(out.shape, dtype=torch.bfloat16, device=dev)
        lo = torch.empty(out.shape, dtype=torch.bfloat16, device=dev)
        pa, ea = lib.ptr_array(planes), lib.ptr_array(effects)
        lut = torch.empty(320 * 32, dtype=torch.float32, device=dev)       # scratch: combined table of the categorical planes
        lib.call('dsb_spatial_stem_fwd', pa, ea, project, ex, ey, entity_num, w2, bias, lut, out, hi, lo, out_c, N, E, H, W)
        ctx.save_for_backward(project, w2, bias, ex, ey, entity_num, *planes, *effects)
        ctx.dims = (N, E, H, W, out_c, tuple(weight.shape))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(hi, lo)
        return out, hi, lo

    @staticmethod
    def backward(ctx, gout, _ghi, _glo):
        project, w2, bias, ex, ey, entity_num = ctx.saved_tensors[:6]
        planes, effects = list(ctx.saved_tensors[6:13]), list(ctx.saved_tensors[13:19])
        N, E, H, W, out_c, wshape = ctx.dims
        if gout is None:
            return (None,) * 20
        dev = gout.device
        gw = torch.zeros((32, 56), dtype=torch.float32, device=dev)
        gb = torch.zeros(32, dtype=torch.float32, device=dev)
        gp = torch.zeros((N, E, 32), dtype=torch.float32, device=dev)
        pa, ea = lib.ptr_array(planes), lib.ptr_array(effects)
        lut = torch.empty(320 * 32, dtype=torch.float32, device=dev)
        lib.call('dsb_spatial_stem_bwd', pa, ea, project, ex, ey, entity_num, w2, bias, lut, gout.contiguous(), out_c, gw, gb, gp,
                 N, E, H, W)
        return (gp, gw.view(wshape), gb, None, None, None, None) + (None,) * 13


def spatial_stem(spatial_info, project, ex, ey, entity_num, weight, bias, out_c: int = 64):
    """relu(project_conv(cat[planes, effects, scatter_connection(project)])) max-pooled 2x2, channels-last
    [N, H/2, W/2, out_c] (first 32 channels real), with the bf16 split attached for the following 3x3 conv."""
    planes = [spatial_info[k].contiguous() for k in STEM_PLANES]
    effects = [spatial_info[k].contiguous() for k in STEM_EFFECTS]
    if _use_kernel(project):
        out, hi, lo = _SpatialStem.apply(project, weight, bias, ex.contiguous(), ey.contiguous(),
                                         entity_num.to(torch.int64).contiguous(), out_c, *planes, *effects)
        return attach_split(out, hi, lo)
    return _standin('spatial_stem')(planes, effects, project, ex, ey, entity_num, weight, bias, out_c)


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
    return _standin('return_scan')(reward, value, rho, gamma_td, lambda_td)


# ------------------------------------------------------------------------------------------------
# per-row categorical statistics (K17) and sampling (K15)
# ------------------------------------------------------------------------------------------------
class _CategoricalStats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, teacher, action, want_mean, flag):
        rows, C = logits.shape
        dev = logits.device
        lse = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        logp = torch.empty(rows, dtype=torch.float32, device=dev)
        ent = torch.empty_like(logp)
        kl = torch.empty_like(logp) if teacher is not None else None
        lse_t = torch.empty_like(lse) if teacher is not None else None
        mean_lp = torch.empty_like(logp) if want_mean else None
        lib.call('dsb_categorical_stats_fwd', logits, teacher, action, lse, logp, ent, kl, lse_t, mean_lp, flag, rows, C)
        ctx.save_for_backward(logits, teacher, action, lse, ent, lse_t)
        ctx.mark_non_differentiable(lse)
        ctx.set_materialize_grads(False)
        if teacher is None:
            kl = torch.zeros_like(logp)
        if mean_lp is None:
            mean_lp = torch.zeros_like(logp)
        return logp, ent, kl, mean_lp, lse

    @staticmethod
    def backward(ctx, g_logp, g_ent, g_kl, g_mean, _g_lse):
        logits, teacher, action, lse, ent, lse_t = ctx.saved_tensors
        rows, C = logits.shape
        grad = torch.empty_like(logits)
        cont = lambda g: None if g is None else g.contiguous().float()
        lib.call('dsb_categorical_stats_bwd', logits, teacher, action, lse, ent, lse_t, cont(g_logp), cont(g_ent),
                 cont(g_kl) if teacher is not None else None, cont(g_mean), grad, rows, C)
        return grad, None, None, None, None


def categorical_stats(logits: torch.Tensor, action: torch.Tensor, teacher: Optional[torch.Tensor] = None,
                      want_mean: bool = False, flag: Optional[torch.Tensor] = None):
    """Per row of ``logits[..., C]``: (log p(action), entropy, KL(teacher || target)); all differentiable wrt logits.
    With want_mean a fourth result, mean_j log p_j (the label-smoothing term of sl_loss.py:16-34), is returned.
    ``flag`` (device int32[1]): bit 2 is set when a label lies outside [0, C) (torch raises there; the kernel clamps).

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
        logp, ent, kl, mean_lp, _ = _CategoricalStats.apply(z, t.contiguous() if t is not None else None, a.contiguous(),
                                                            want_mean, flag)
    else:
        logp, ent, kl, mean_lp = _standin('categorical_stats')(z, t, a, flag)
    if want_mean:
        return logp.view(shape), ent.view(shape), kl.view(shape), mean_lp.view(shape)
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
    return _standin('sample_categorical')(logits, q)


# ------------------------------------------------------------------------------------------------
# selected-units pointer network, sampling path (K12)
# ------------------------------------------------------------------------------------------------
def su_sample(weights16, emb0, key, valid_mask, entity_num, su_mask, temperature: float, rng: str = 'cuda',
              max_steps: int = 64, poll: bool = True):
    """The sampling loop of SelectedUnitsHead._query (action_arg_head.py:262-314) with ONE kernel per step.

    key [N,S,32] (end token at slot entity_num), valid_mask [N,S] bool (slots < entity_num + 1), su_mask [N] bool (rows
    whose action type takes a unit selection).  rng='cpu' draws every step's Exp(1) variates from the CPU generator in the
    reference's order and checks the all-ended condition after every step (so the stream is consumed identically);
    rng='cuda' draws on device and polls the exit condition every 8 steps (extra steps do not change any output).
    Returns (logits [N,steps,S], units [N,steps], ae [N,1024], selected_units_num [N])."""
    N, S, _ = key.shape
    dev = key.device
    # end flag is not available at the first selection (pure device ops: an indexed assignment of a Python scalar would be
    # a host -> device copy, which CUDA-graph capture forbids)
    mask = valid_mask & (torch.arange(S, device=dev).unsqueeze(0) != entity_num.unsqueeze(1))
    mask = mask.to(torch.uint8).contiguous()
    ae = emb0.detach().clone().contiguous()
    emb0c = emb0.detach().contiguous()
    h = torch.zeros((N, 32), dtype=torch.float32, device=dev)
    c = torch.zeros((N, 32), dtype=torch.float32, device=dev)
    ksum = torch.zeros((N, 32), dtype=torch.float32, device=dev)
    count = torch.zeros(N, dtype=torch.int32, device=dev)
    end_flag = (~su_mask).to(torch.uint8).contiguous()
    num = torch.where(su_mask, torch.full((N,), max_steps, dtype=torch.int64, device=dev),
                      torch.zeros(N, dtype=torch.int64, device=dev)).contiguous()
    logits = torch.empty((max_steps, N, S), dtype=torch.float32, device=dev)
    results = torch.empty((max_steps, N), dtype=torch.int64, device=dev)
    ended = torch.ones(max_steps, dtype=torch.int32, device=dev)
    en = entity_num.to(torch.int64).contiguous()
    keyc = key.detach().contiguous()
    keep = [w.detach().contiguous() for w in weights16]      # keep the contiguous copies alive during the launches
    warr = lib.ptr_array(keep)
    steps = max_steps
    for i in range(max_steps):
        if rng == 'cpu':
            q = torch.empty((N, S), dtype=torch.float32).exponential_(1).to(dev)
        else:
            q = torch.empty((N, S), dtype=torch.float32, device=dev).exponential_(1)
        lib.call('dsb_su_sample_step', warr, emb0c, ae, keyc, h, c, mask, ksum, count, end_flag, num, en,
                 results[i - 1] if i > 0 else None, q, logits[i], results[i], ended[i:i + 1], N, S, i, float(temperature))
        if rng == 'cpu':
            if int(ended[i].item()) != 0:
                steps = i + 1
                break
        elif poll and ((i & 7) == 7 or i == max_steps - 1):
            # (poll=False: fixed step count, no host read - the form a CUDA graph can capture; the caller trims)
            flags = ended[:i + 1].tolist()
            if 1 in flags:
                steps = flags.index(1) + 1
                break
    return logits[:steps].transpose(0, 1).contiguous(), results[:steps].transpose(0, 1).contiguous(), ae, num


# ------------------------------------------------------------------------------------------------
# split-precision tcgen05 linear (K2/K3 workhorse)
# ------------------------------------------------------------------------------------------------
def attach_split(t: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor) -> torch.Tensor:
    """Remember the bf16 (hi, lo) pair a producer kernel already wrote for `t` so the consuming GEMM skips its split."""
    t._dsb_split = (hi, lo)
    return t


_PLACEHOLDER = {}

# ---- derived forms of a weight (bf16 pair, conv GEMM matrices) are rebuilt at most once per optimiser step ----------------
# A learner step calls every encoder layer once per chunk (and again in backward); re-splitting the same unchanged weight
# each time was ~1000 tiny launches per step.  The cache lives on the parameter object and is keyed by the tensor version
# (bumped by every in-place torch op: load_state_dict, manual edits) and by WEIGHT_EPOCH, which FlatAdam bumps because its
# kernel updates the arena behind torch's back.
WEIGHT_EPOCH = [0]


def weight_cached(w: torch.Tensor, key: str, build):
    if not (w.is_leaf and w.is_cuda):
        return build()
    cache = w.__dict__.setdefault('_dsb_wcache', {})
    stamp = (w._version, WEIGHT_EPOCH[0], w.data_ptr())
    hit = cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    val = build()
    cache[key] = (stamp, val, build)
    return val


def _flat_tensors(v):
    if torch.is_tensor(v):
        return [v]
    if isinstance(v, (tuple, list)):
        return [t for x in v for t in _flat_tensors(x)]
    return []


def refresh_weight_cache(params) -> int:
    """Recompute every cached derived weight form of `params` INTO THE SAME BUFFERS.  A captured CUDA graph (serving.py) has
    the addresses of those buffers baked into its launches, so after new weights arrive they must be rewritten in place rather
    than replaced.  Returns the number of refreshed entries."""
    n = 0
    for p in params:
        cache = p.__dict__.get('_dsb_wcache')
        if not cache:
            continue
        for key, entry in list(cache.items()):
            if len(entry) < 3:
                continue
            _, old, build = entry
            for dst, src in zip(_flat_tensors(old), _flat_tensors(build())):
                dst.copy_(src)
            cache[key] = ((p._version, WEIGHT_EPOCH[0], p.data_ptr()), old, build)
            n += 1
    return n


def invalidate_weight_cache() -> None:
    """Call after changing parameter storage through an alias torch's version counter does not see (the flat arena,
    ``p.data``): Model.load_state_dict / .to() / DistModule.broadcast_params and FlatAdam.step do."""
    WEIGHT_EPOCH[0] += 1


def weight_split(w: torch.Tensor):
    """bf16 (hi, lo) pair of a parameter, cached per optimiser step."""
    return weight_cached(w, 'split', lambda: split_bf16(w.detach()))



def pair_only_placeholder(shape, device) -> torch.Tensor:
    """Stand-in for the fp32 output of a GEMM launched with emit_split='only': a stride-0 view of one NaN, so it costs no
    memory or bandwidth and anything that wrongly reads it (instead of the attached bf16 pair) poisons the result loudly."""
    t = _PLACEHOLDER.get(device)
    if t is None:
        t = _PLACEHOLDER[device] = torch.full((), float('nan'), dtype=torch.float32, device=device)
    return t.expand(shape)


def split_bf16(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo == x to ~2^-17 relative."""
    cached = getattr(x, '_dsb_split', None)
    if cached is not None and cached[0].shape == x.shape:
        return cached
    x = x.contiguous()
    if _use_kernel(x):
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lib.call('dsb_split_bf16', x, hi, lo, x.numel())
        return hi, lo
    return _standin('split_bf16')(x)


# DEBUG ONLY: DSB_DISABLE_TCGEN05=1 routes every fc_block through the library matmul so the rest of the pipeline
# can be bisected on a GPU box when the tensor-core kernel is under suspicion.  Never set in tests of record/bench.
_TCGEN05_OFF = os.environ.get('DSB_DISABLE_TCGEN05', '0') == '1'


def gemm_eligible(N: int, K: int) -> bool:
    return (not _TCGEN05_OFF) and N % 128 == 0 and K % 64 == 0


def gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu: bool, terms: int = 3, want_split: bool = False):
    """C = act(A @ W^T + bias) on the tensor cores; A [M,K], W [N,K] as bf16 (hi, lo) pairs -> fp32 [M,N].
    a_lo None: A is exact in bf16 (two products instead of three)."""
    M, K = a_hi.shape
    N = w_hi.shape[0]
    assert w_hi.shape[1] == K and N % 64 == 0 and K % 64 == 0, (M, N, K)
    if a_lo is None and a_hi.is_cuda:
        c = torch.empty((M, N), dtype=torch.float32, device=a_hi.device) if want_split != 'only' else None
        c_hi = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        c_lo = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        _gemm_ex(a_hi=a_hi, a_lo=None, b_hi=w_hi, b_lo=w_lo, bias=bias, alpha=1.0, relu=1 if relu else 0, terms=terms, c=c,
                 c_hi=c_hi, c_lo=c_lo, m=M, n=N, k=K, batch=1, inner=1, splits=1, a_exact=1)
        if c is None:
            c = pair_only_placeholder((M, N), a_hi.device)
        return (c, c_hi, c_lo) if want_split else c
    if _use_kernel(a_hi):
        c = torch.empty((M, N), dtype=torch.float32, device=a_hi.device) if want_split != 'only' else None
        c_hi = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        c_lo = torch.empty((M, N), dtype=torch.bfloat16, device=a_hi.device) if want_split else None
        lib.call('dsb_gemm_bf16_split', a_hi, a_lo, w_hi, w_lo, bias, c, c_hi, c_lo, M, N, K, terms, 1 if relu else 0)
        if c is None:
            c = pair_only_placeholder((M, N), a_hi.device)
        return (c, c_hi, c_lo) if want_split else c
    c = _standin('gemm_split')(a_hi, a_lo, w_hi, w_lo, bias, relu, terms)
    if want_split:
        h, l = split_bf16(c)
        return c, h, l
    return c


def _gemm_ex(**kw) -> None:
    lib.gemm_ex(**kw)


def _skinny_gemm(a_hi, a_lo, w_hi, w_lo, bias, relu: bool, terms: int):
    """Few rows, very long reduction (the spatial encoder's fc: 264 x 32768 -> 256 gives 6 output tiles for 148 SMs): split
    the reduction over CTAs (partials summed by TMA reduce-add) and apply bias / ReLU afterwards.  None when not worth it."""
    M, K = a_hi.shape
    N = w_hi.shape[0]
    if not (a_hi.is_cuda and K >= 4096 and N % 128 == 0 and K % 64 == 0):
        return None
    tiles = ((M + 127) // 128) * (N // 128)
    if tiles > 37:
        return None
    splits = _pick_splits(tiles, K)
    if splits < 2:
        return None
    c = torch.zeros((M, N), dtype=torch.float32, device=a_hi.device)
    _gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, alpha=1.0, terms=terms, c=c, m=M, n=N, k=K, batch=1, inner=1,
             splits=splits, c_row_split=0, c_accumulate=1, bn=128)
    if bias is not None:
        c += bias
    return torch.relu_(c) if relu else c


def relu_bwd_split(gy2: torch.Tensor, y: Optional[torch.Tensor], need_bias: bool, need_g: bool = False,
                   need_split: bool = True, bias_grad_out: Optional[torch.Tensor] = None):
    """(g_hi, g_lo, bias_grad or None, g or None) with g = gy * (y > 0): one pass instead of compare + mul + split + sum.
    With ``bias_grad_out`` ([N] fp32, e.g. the bias parameter's .grad) the column sums are added into it by the kernel and
    no bias gradient is returned."""
    rows, N = gy2.shape
    dev = gy2.device
    hi = torch.empty((rows, N), dtype=torch.bfloat16, device=dev) if need_split else None
    lo = torch.empty((rows, N), dtype=torch.bfloat16, device=dev) if need_split else None
    g = torch.empty((rows, N), dtype=torch.float32, device=dev) if need_g else None
    is_bf16 = 1 if (y is not None and y.dtype == torch.bfloat16) else 0
    if bias_grad_out is not None:
        assert bias_grad_out.numel() == N and bias_grad_out.is_contiguous()
        lib.call('dsb_relu_bwd_split', gy2, y, is_bf16, g, hi, lo, bias_grad_out, 1, rows, N)
        return hi, lo, None, g
    blocks = lib.load().dsb_relu_bwd_split_blocks(rows, N)
    cs = torch.empty((blocks, N), dtype=torch.float32, device=dev) if need_bias else None
    lib.call('dsb_relu_bwd_split', gy2, y, is_bf16, g, hi, lo, cs, 0, rows, N)
    return hi, lo, (cs.sum(0) if need_bias else None), g


# Parameters of the Model live in one flat arena and their .grad is a view of the flat gradient arena (model.py).  When a
# Function's weight / bias is such a leaf, its backward adds the gradient straight into .grad (TMA reduce-add in the dW GEMM
# epilogue, atomics for the bias column sums) and hands autograd None: no zero-filled temporary, no AccumulateGrad add
# kernel per parameter per encoder chunk.
def _grad_slot(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if p is None or not p.is_leaf or not p.is_cuda:
        return None
    g = p.grad
    if g is None or g.shape != p.shape or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() % 16 != 0:
        return None
    return g


def _pick_splits(tiles: int, k: int) -> int:
    """split-K factor for a reduction of length k (tokens) over `tiles` output tiles: fill ~148 SMs, keep >= 512 k per
    split, and k % (64 * splits) == 0."""
    s = 1
    while tiles * s * 2 <= 296 and k % (64 * s * 2) == 0 and k // (s * 2) >= 512:
        s *= 2
    return s


def weight_grad(g_hi, g_lo, x_hi, x_lo, terms: int = 3, accumulate_into: Optional[torch.Tensor] = None):
    """dW[N,K] = dY^T X with both operands read in place (MN-major: the reduction runs over rows = tokens), split-K over
    tokens on the tensor cores, partial tiles summed by the copy engine (TMA reduce-add).  With ``accumulate_into`` the
    result is added to that [N,K] tensor (the parameter's .grad) and None is returned."""
    M, N = g_hi.shape
    K = x_hi.shape[1]
    kk = _pad_to(M, 64)                                 # rows behind M are TMA zero fill on both operands (any token count)
    splits = _pick_splits(((N + 127) // 128) * (K // 128), kk)
    bx = 1 if x_lo is None else 0                       # the activation is exact in bf16: no dY_hi x X_lo product
    if accumulate_into is not None:
        _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=x_hi, b_lo=x_lo, a_mn=1, b_mn=1, alpha=1.0, terms=terms, c=accumulate_into, m=N,
                 n=K, k=kk, batch=1, inner=1, splits=splits, c_row_split=0, c_accumulate=1, b_exact=bx)
        return None
    gw = torch.zeros((N, K), dtype=torch.float32, device=g_hi.device) if splits > 1 else \
        torch.empty((N, K), dtype=torch.float32, device=g_hi.device)
    _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=x_hi, b_lo=x_lo, a_mn=1, b_mn=1, alpha=1.0, terms=terms, c=gw, m=N, n=K, k=kk,
             batch=1, inner=1, splits=splits, c_row_split=0, c_accumulate=1 if splits > 1 else 0, b_exact=bx)
    return gw


class _SplitLinear(torch.autograd.Function):
    """y = act(x W^T + b): forward, input gradient (dY . W, W read MN-major in place) and weight gradient
    (dY^T . X, split-K over tokens) all on the tcgen05 kernel with 3-term split products."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, terms, x_hi=None, x_lo=None, emit_split=False, fork=False):
        x2 = x.reshape(-1, x.shape[-1])
        if x_hi is not None:
            a_hi, a_lo = x_hi.reshape(x2.shape), (x_lo.reshape(x2.shape) if x_lo is not None else None)
        else:
            a_hi, a_lo = split_bf16(x2)
        w_hi, w_lo = weight_split(weight)
        oshape = (*x.shape[:-1], weight.shape[0])
        ctx.set_materialize_grads(False)      # no zero-filled bf16 'gradients' for the (hi, lo) side outputs
        if emit_split:
            y, y_hi, y_lo = gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu, terms, want_split=emit_split)
        else:
            y = _skinny_gemm(a_hi, a_lo, w_hi, w_lo, bias, relu, terms) if a_lo is not None else None
            if y is None:
                y = gemm_split(a_hi, a_lo, w_hi, w_lo, bias, relu, terms)
        # the ReLU mask only needs the sign: when the bf16 pair is emitted (and kept by the consumer anyway) save its hi
        # half instead of the fp32 output
        ctx.save_for_backward(a_hi, a_lo, w_hi, w_lo, (y_hi if emit_split else y) if relu else None)
        ctx.relu, ctx.terms, ctx.has_bias, ctx.xshape = relu, terms, bias is not None, x.shape
        ctx.weight_ref, ctx.bias_ref = weight, bias
        # fork: also hand the input back as a second differentiable output (the residual branch of a transformer sub-layer
        # takes it from here), so both gradients of x meet INSIDE this node and the dX GEMM adds the branch's gradient in its
        # epilogue instead of autograd launching a separate add over [tokens, d]
        x_pass = x.view_as(x) if fork else None
        if emit_split:
            y_hi, y_lo = y_hi.view(oshape), y_lo.view(oshape)
            ctx.mark_non_differentiable(y_hi, y_lo)
            return y.view(oshape), y_hi, y_lo, x_pass
        return y.view(oshape), None, None, x_pass

    @staticmethod
    def backward(ctx, gy, _ghi=None, _glo=None, g_pass=None):
        a_hi, a_lo, w_hi, w_lo, y = ctx.saved_tensors
        if gy is None:
            return (g_pass,) + (None,) * 8
        pre = getattr(gy, '_dsb_grad_pair', None)
        K = a_hi.shape[1]
        gx = gw = gb = None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if pre is not None and not ctx.relu:
            # the producer of this gradient (attention backward, ...) already wrote it as the bf16 pair the GEMMs below read
            g_hi, g_lo, bias_done = pre
            M, N = g_hi.shape
            on_gpu, gy2, g = True, None, None
            if want_b and not bias_done:
                slot = _grad_slot(ctx.bias_ref)
                acc = slot if slot is not None else torch.zeros(N, dtype=torch.float32, device=g_hi.device)
                lib.call('dsb_colsum_pair', g_hi, g_lo, acc, M, N)
                gb = None if slot is not None else acc
        else:
            gy2 = gy.reshape(-1, gy.shape[-1]).contiguous()
            M, N = gy2.shape
            on_gpu = gy2.is_cuda
        if pre is not None and not ctx.relu:
            pass
        elif on_gpu and N % 4 == 0:
            g_hi, g_lo, gb, g = relu_bwd_split(gy2, y if ctx.relu else None, want_b, need_g=False,
                                               bias_grad_out=_grad_slot(ctx.bias_ref) if want_b else None)
        else:
            g = gy2 * (y.reshape(gy2.shape) > 0) if ctx.relu else gy2
            g_hi, g_lo = split_bf16(g) if on_gpu else (None, None)
            gb = g.sum(0) if want_b else None
        if ctx.needs_input_grad[0]:
            if on_gpu and K % 128 == 0 and N % 64 == 0:
                gx = torch.empty((M, K), dtype=torch.float32, device=g_hi.device)
                res = g_pass.reshape(M, K).contiguous() if g_pass is not None else None
                _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=w_hi, b_lo=w_lo, b_mn=1, alpha=1.0, terms=ctx.terms, c=gx, m=M, n=K,
                         k=N, batch=1, inner=1, splits=1, residual=res)
            else:
                gfull = g if g is not None else (g_hi.float() + g_lo.float())
                gx = gfull @ (w_hi.float() + w_lo.float())
                if g_pass is not None:
                    gx = gx + g_pass.reshape(gx.shape)
            gx = gx.view(ctx.xshape)
        elif g_pass is not None:
            gx = g_pass
        if ctx.needs_input_grad[1]:
            if on_gpu and N % 64 == 0 and K % 128 == 0:
                gw = weight_grad(g_hi, g_lo, a_hi, a_lo, ctx.terms, accumulate_into=_grad_slot(ctx.weight_ref))
            else:
                gfull = g if g is not None else (g_hi.float() + g_lo.float())
                gw = gfull.t() @ (a_hi.float() + (a_lo.float() if a_lo is not None else 0))
        return gx, gw, gb, None, None, None, None, None, None


class _FFN(torch.autograd.Function):
    """m = relu(relu(x W1^T + b1) W2^T + b2): the transformer MLP (module_utils.py:130-139) as one node that does NOT keep its
    hidden activation h [tokens, 4d] for backward - h is the largest tensor a transformer layer saves (as large as the
    attention probabilities) and costs one 3-term GEMM to rebuild, far cheaper per byte than re-running the whole encoder
    chunk, which is what running out of HBM otherwise forces (model.py: keep_chunks)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, terms, x_hi, x_lo, fork=False):
        K = x.shape[-1]
        a_hi, a_lo = x_hi.reshape(-1, K), x_lo.reshape(-1, K)
        w1_hi, w1_lo = weight_split(w1)
        w2_hi, w2_lo = weight_split(w2)
        _, h_hi, h_lo = gemm_split(a_hi, a_lo, w1_hi, w1_lo, b1, True, terms, want_split='only')
        m = gemm_split(h_hi, h_lo, w2_hi, w2_lo, b2, True, terms)
        ctx.save_for_backward(a_hi, a_lo, w1_hi, w1_lo, w2_hi, w2_lo, m)
        ctx.refs = (w1, b1, w2, b2)
        ctx.terms, ctx.xshape = terms, x.shape
        ctx.set_materialize_grads(False)
        return m.view(*x.shape[:-1], w2.shape[0]), (x.view_as(x) if fork else None)

    @staticmethod
    def backward(ctx, gy, g_pass=None):
        a_hi, a_lo, w1_hi, w1_lo, w2_hi, w2_lo, m = ctx.saved_tensors
        w1, b1, w2, b2 = ctx.refs
        terms = ctx.terms
        if gy is None:
            return (g_pass,) + (None,) * 8
        gy2 = gy.reshape(m.shape).contiguous()
        M = gy2.shape[0]
        H, K = w1_hi.shape
        dev = gy2.device
        # second layer
        g2_hi, g2_lo, gb2, _ = relu_bwd_split(gy2, m, ctx.needs_input_grad[4], bias_grad_out=_grad_slot(b2))
        _, h_hi, h_lo = gemm_split(a_hi, a_lo, w1_hi, w1_lo, b1, True, terms, want_split='only')      # rebuild h
        gw2 = weight_grad(g2_hi, g2_lo, h_hi, h_lo, terms, accumulate_into=_grad_slot(w2)) if ctx.needs_input_grad[3] else None
        # dh = g2 W2 through the first layer's ReLU: the mask (sign of the rebuilt bf16 hidden) is applied in the GEMM epilogue and
        # only the (hi, lo) pair the next two GEMMs read is written - the fp32 dh and its separate masking pass do not exist
        g1_hi = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
        g1_lo = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
        # (and the column sums of the masked gradient = the first layer's bias gradient are formed in the same epilogue)
        gb1 = acc1 = None
        if ctx.needs_input_grad[2]:
            slot1 = _grad_slot(b1)
            acc1 = slot1 if slot1 is not None else torch.zeros(H, dtype=torch.float32, device=dev)
            gb1 = None if slot1 is not None else acc1
        _gemm_ex(a_hi=g2_hi, a_lo=g2_lo, b_hi=w2_hi, b_lo=w2_lo, b_mn=1, alpha=1.0, terms=terms, c=None, c_hi=g1_hi, c_lo=g1_lo,
                 m=M, n=H, k=w2_hi.shape[0], batch=1, inner=1, splits=1, relu_mask=h_hi, colsum=acc1)
        del g2_hi, g2_lo
        # first layer
        del h_hi, h_lo
        gw1 = weight_grad(g1_hi, g1_lo, a_hi, a_lo, terms, accumulate_into=_grad_slot(w1)) if ctx.needs_input_grad[1] else None
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty((M, K), dtype=torch.float32, device=dev)
            res = g_pass.reshape(M, K).contiguous() if g_pass is not None else None      # the residual branch's gradient
            _gemm_ex(a_hi=g1_hi, a_lo=g1_lo, b_hi=w1_hi, b_lo=w1_lo, b_mn=1, alpha=1.0, terms=terms, c=gx, m=M, n=K, k=H, batch=1,
                     inner=1, splits=1, residual=res)
            gx = gx.view(ctx.xshape)
        elif g_pass is not None:
            gx = g_pass
        return gx, gw1, gb1, gw2, gb2, None, None, None, None


def ffn(x: torch.Tensor, w1, b1, w2, b2, terms: int = 3, fork: bool = False):
    """relu(fc2(relu(fc1(x)))) of a transformer layer; on the GPU (x carrying its bf16 pair) through _FFN.  fork=True returns
    (m, x') with x' the input handed back for the residual branch (see _SplitLinear.forward)."""
    sp = getattr(x, '_dsb_split', None)
    H, K = w1.shape
    if (x.is_cuda and sp is not None and sp[0].shape == x.shape and gemm_eligible(H, K) and gemm_eligible(w2.shape[0], H)
            and K % 128 == 0 and w2.shape[0] % 128 == 0 and (x.numel() // K) % 64 == 0 and x.numel() // K >= 128):
        m, x_pass = _FFN.apply(x, w1, b1, w2, b2, terms, sp[0], sp[1], fork)
        return (m, x_pass) if fork else m
    m = linear(linear(x, w1, b1, True, terms, emit_split='only' if x.is_cuda else False), w2, b2, True, terms)
    return (m, x) if fork else m


class _EntityAttention(torch.autograd.Function):
    """softmax(Q K^T / sqrt(d) + key mask) V for all (observation, head) pairs of the entity transformer
    (model/module_utils.py:95-110), forward and backward, as batched tcgen05 GEMMs that address Q, K, V inside the
    QKV activation by coordinates (no split / permute / contiguous copies) plus two row-softmax kernels."""

    @staticmethod
    def forward(ctx, qkv, entity_num, heads, hd, q_hi=None, q_lo=None, qkv_bias=None, pair_grad=False):
        ctx.in_shape = qkv.shape
        ctx.bias_ref = qkv_bias
        ctx.pair_grad = pair_grad
        qkv = qkv.reshape(-1, qkv.shape[-1])
        NS, W3 = qkv.shape
        n_obs = entity_num.shape[0]
        S = NS // n_obs
        H, D = heads, hd
        dev = qkv.device
        if q_hi is None:                       # otherwise the producer GEMM already wrote the pair (qkv may be a placeholder)
            q_hi, q_lo = split_bf16(qkv)
        scores = torch.empty((n_obs * H * S, S), dtype=torch.float32, device=dev)
        _gemm_ex(a_hi=q_hi, a_lo=q_lo, b_hi=q_hi, b_lo=q_lo, alpha=1.0 / math.sqrt(D), terms=3, c=scores, m=S, n=S, k=D,
                 batch=n_obs * H, inner=H, splits=1, a_col_base=0, a_col_inner=D, a_row_outer=S,
                 b_col_base=H * D, b_col_inner=D, b_row_outer=S, c_row_outer=H * S, c_row_inner=S)
        p_hi = torch.empty((n_obs * H * S, S), dtype=torch.bfloat16, device=dev)
        p_lo = torch.empty_like(p_hi)
        lib.call('dsb_attn_softmax_fwd', scores, entity_num, H * S, p_hi, p_lo, n_obs * H * S, S)
        del scores
        # the context only feeds the projection GEMM: write it as a bf16 pair, no fp32 copy
        o_hi = torch.empty((NS, H * D), dtype=torch.bfloat16, device=dev)
        o_lo = torch.empty_like(o_hi)
        _gemm_ex(a_hi=p_hi, a_lo=p_lo, b_hi=q_hi, b_lo=q_lo, b_mn=1, alpha=1.0, terms=3, c=None, c_hi=o_hi, c_lo=o_lo, m=S,
                 n=D, k=S, batch=n_obs * H, inner=H, splits=1, a_row_outer=H * S, a_row_inner=S,
                 b_col_base=2 * H * D, b_col_inner=D, b_row_outer=S, c_row_outer=S, c_col_inner=D)
        ctx.save_for_backward(q_hi, q_lo, p_hi, p_lo, entity_num)
        ctx.dims = (n_obs, S, H, D)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(o_hi, o_lo)
        return pair_only_placeholder((NS, H * D), dev), o_hi, o_lo

    @staticmethod
    def backward(ctx, g_out, _ghi=None, _glo=None):
        q_hi, q_lo, p_hi, p_lo, entity_num = ctx.saved_tensors
        n_obs, S, H, D = ctx.dims
        dev = g_out.device
        alpha = 1.0 / math.sqrt(D)
        g_hi, g_lo = split_bf16(g_out.contiguous())
        # the gradient of the QKV activation only feeds the QKV layer's dX / dW GEMMs: the three products below write it as the
        # bf16 pair those GEMMs read (no fp32 tensor, no separate split pass) and add its column sums - the QKV bias gradient -
        # from their epilogues
        # (only when the consumer is known to be a tensor-core linear that reads the pair: ctx.pair_grad; otherwise plain fp32)
        bslot = None
        if ctx.pair_grad:
            dq_hi = torch.empty((n_obs * S, 3 * H * D), dtype=torch.bfloat16, device=dev)
            dq_lo = torch.empty_like(dq_hi)
            bslot = _grad_slot(ctx.bias_ref) if ctx.bias_ref is not None else None
            out = dict(c=None, c_hi=dq_hi, c_lo=dq_lo, colsum=bslot)
        else:
            dq32 = torch.empty((n_obs * S, 3 * H * D), dtype=torch.float32, device=dev)
            out = dict(c=dq32)
        common = dict(terms=3, batch=n_obs * H, inner=H, splits=1)
        # dP[q,k] = sum_d dO[q,d] V[k,d]
        dp = torch.empty((n_obs * H * S, S), dtype=torch.float32, device=dev)
        _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=q_hi, b_lo=q_lo, alpha=1.0, c=dp, m=S, n=S, k=D,
                 a_col_inner=D, a_row_outer=S, b_col_base=2 * H * D, b_col_inner=D, b_row_outer=S,
                 c_row_outer=H * S, c_row_inner=S, **common)
        # dV[k,d] = sum_q P[q,k] dO[q,d]          (both operands reduce over rows -> MN-major)
        _gemm_ex(a_hi=p_hi, a_lo=p_lo, b_hi=g_hi, b_lo=g_lo, a_mn=1, b_mn=1, alpha=1.0, m=S, n=D, k=S,
                 a_row_outer=H * S, a_row_inner=S, b_col_inner=D, b_row_outer=S,
                 c_row_outer=S, c_col_base=2 * H * D, c_col_inner=D, **out, **common)
        ds_hi = torch.empty_like(p_hi)
        ds_lo = torch.empty_like(p_lo)
        lib.call('dsb_attn_softmax_bwd', p_hi, p_lo, dp, entity_num, H * S, ds_hi, ds_lo, n_obs * H * S, S)
        del dp
        # dQ[q,d] = alpha * sum_k dS[q,k] K[k,d]
        _gemm_ex(a_hi=ds_hi, a_lo=ds_lo, b_hi=q_hi, b_lo=q_lo, b_mn=1, alpha=alpha, m=S, n=D, k=S,
                 a_row_outer=H * S, a_row_inner=S, b_col_base=H * D, b_col_inner=D, b_row_outer=S,
                 c_row_outer=S, c_col_base=0, c_col_inner=D, **out, **common)
        # dK[k,d] = alpha * sum_q dS[q,k] Q[q,d]
        _gemm_ex(a_hi=ds_hi, a_lo=ds_lo, b_hi=q_hi, b_lo=q_lo, a_mn=1, b_mn=1, alpha=alpha, m=S, n=D, k=S,
                 a_row_outer=H * S, a_row_inner=S, b_col_base=0, b_col_inner=D, b_row_outer=S,
                 c_row_outer=S, c_col_base=H * D, c_col_inner=D, **out, **common)
        if not ctx.pair_grad:
            return dq32.view(ctx.in_shape), None, None, None, None, None, None, None
        dqkv = pair_only_placeholder(ctx.in_shape, dev)
        dqkv._dsb_grad_pair = (dq_hi, dq_lo, bslot is not None)       # (hi, lo, bias gradient already accumulated)
        return dqkv, None, None, None, None, None, None, None


def entity_attention(qkv: torch.Tensor, entity_num: torch.Tensor, heads: int, hd: int,
                     qkv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [N, S, 3*heads*hd] (q | k | v, each head-major) -> context [N, S, heads*hd]; keys >= entity_num masked.
    On the GPU the context is returned as a pair-only tensor (fp32 placeholder + attached bf16 (hi, lo) pair, see
    pair_only_placeholder): its one consumer is the projection GEMM."""
    N, S, W3 = qkv.shape
    if _use_kernel(qkv) and S % 128 == 0 and hd % 128 == 0:
        sp = getattr(qkv, '_dsb_split', None)
        # (qkv goes in un-reshaped: no view node between this Function and the QKV linear, so the gradient object this backward
        # returns - a placeholder carrying the bf16 pair - reaches the linear's backward as is)
        if sp is not None and sp[0].shape == qkv.shape:
            out, o_hi, o_lo = _EntityAttention.apply(qkv, entity_num.to(torch.int64).contiguous(), heads,
                                                     hd, sp[0].reshape(N * S, W3), sp[1].reshape(N * S, W3), qkv_bias,
                                                     bool(getattr(qkv, '_dsb_reads_grad_pair', False)))
        else:
            out, o_hi, o_lo = _EntityAttention.apply(qkv.contiguous(), entity_num.to(torch.int64).contiguous(), heads, hd,
                                                     None, None, qkv_bias, False)
        shape = (N, S, heads * hd)
        return attach_split(out.view(shape), o_hi.view(shape), o_lo.view(shape))
    q, k, v = qkv.view(N, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    score = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    if entity_num is not None:
        key_mask = torch.arange(S, device=qkv.device).unsqueeze(0) < entity_num.unsqueeze(1)
        score = score.masked_fill(~key_mask.view(N, 1, 1, S), -1e9)
    return torch.matmul(torch.softmax(score, dim=-1), v).permute(0, 2, 1, 3).reshape(N, S, heads * hd)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           terms: int = 3, emit_split: bool = False, allow_n64: bool = False, fork: bool = False, exact_input: bool = False):
    """fc_block forward (ctools/torch_utils/network/nn_module.py:231-270): tcgen05 split GEMM when the shape
    tiles (N % 128 == 0, K % 64 == 0), plain library matmul for the odd small layers."""
    N, K = weight.shape
    elig = gemm_eligible(N, K) or (allow_n64 and not _TCGEN05_OFF and N % 64 == 0 and K % 128 == 0)   # 64-wide tiles
    elig = elig and x.dtype == torch.float32
    if elig and (x.is_cuda or _HOST_LOGIC_TESTING) and x.numel() // K >= 1:
        sp = getattr(x, '_dsb_split', None)
        if sp is None or sp[0].shape != x.shape:
            sp = (None, None)
        emit = emit_split if x.is_cuda else False       # True: fp32 + pair, 'only': pair only
        y, y_hi, y_lo, x_pass = _SplitLinear.apply(x, weight, bias, relu, terms, sp[0], sp[1], emit, fork and x.is_cuda)
        y = attach_split(y, y_hi, y_lo) if emit else y
        if x.is_cuda and not relu:
            y._dsb_reads_grad_pair = True      # this node's backward accepts its gradient as a bf16 pair (_dsb_grad_pair)
        return (y, x_pass if x_pass is not None else x) if fork else y
    if _use_kernel(x) and not _TCGEN05_OFF and x.numel() > 0:
        y = linear_any(x, weight, bias, relu, terms, exact_input)
        return (y, x) if fork else y
    y = F.linear(x.float(), weight, bias)
    y = torch.relu(y) if relu else y
    return (y, x) if fork else y


# ------------------------------------------------------------------------------------------------
# fc_block for ANY (N, K): operands padded to tile multiples (scalar encoder inputs, head MLPs, value_fc)
# ------------------------------------------------------------------------------------------------
_PACK_DTYPE = {torch.uint8: 0, torch.int16: 1, torch.int8: 2, torch.float16: 3, torch.float32: 4, torch.int64: 5}


def pack_pair(x2: torch.Tensor, Kp: int, want_lo: bool = True):
    """[rows, K] of any wire dtype -> bf16 (hi, lo) [rows, Kp] with zero columns behind K: the A operand of a tcgen05 GEMM whose
    reduction length is not a tile multiple, or whose input is still an integer observation (no .float() copy)."""
    rows, K = x2.shape
    assert x2.is_cuda and x2.dtype in _PACK_DTYPE and x2.stride(1) == 1, (x2.dtype, x2.stride())
    hi = torch.empty((rows, Kp), dtype=torch.bfloat16, device=x2.device)
    lo = torch.empty((rows, Kp), dtype=torch.bfloat16, device=x2.device) if want_lo else None
    # rows may be strided (a column slice of a wider GEMM output): the kernel takes the row pitch, so pass the raw address
    lib.call('dsb_pack_pair', x2.data_ptr(), _PACK_DTYPE[x2.dtype], rows, K, x2.stride(0), hi, lo, Kp)
    return hi, lo


def _padded_weight(weight: torch.Tensor, bias: Optional[torch.Tensor], Np: int, Kp: int):
    w = F.pad(weight.detach(), (0, Kp - weight.shape[1], 0, Np - weight.shape[0]))
    hi, lo = split_bf16(w)
    b = F.pad(bias.detach(), (0, Np - bias.shape[0])).contiguous() if bias is not None else None
    return hi, lo, b


class _PaddedLinear(torch.autograd.Function):
    """y = act(x W^T + b) for shapes the tile grid does not divide: N is padded to a multiple of 64 and K to a multiple of 64
    with zeros (weights once per optimiser step, the activation by the packing kernel); forward, dX and dW all run on the
    tcgen05 kernel.  Returns the PADDED result [rows, Np]; the caller slices (autograd then pads the gradient back)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, terms, exact_input):
        N, K = weight.shape
        Kp = _pad_to(K, 64)
        Np = _pad_to(N, 64) if N <= 64 else _pad_to(N, 128)
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M = x2.shape[0]
        a_hi, a_lo = pack_pair(x2, Kp, want_lo=not exact_input)
        w_hi, w_lo, b_pad = weight_cached(weight, 'padded_%d_%d' % (Np, Kp), lambda: _padded_weight(weight, bias, Np, Kp))
        if bias is not None and not bias.is_leaf:
            b_pad = F.pad(bias.detach(), (0, Np - N)).contiguous()
        y = torch.empty((M, Np), dtype=torch.float32, device=x.device)
        _gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b_pad, alpha=1.0, relu=1 if relu else 0, terms=terms, c=y,
                 m=M, n=Np, k=Kp, batch=1, inner=1, splits=1, a_exact=1 if exact_input else 0)
        ctx.save_for_backward(a_hi, a_lo, w_hi, w_lo, y if relu else None)
        ctx.meta = (N, K, Np, Kp, M, relu, terms, bias is not None, tuple(x.shape), exact_input)
        ctx.weight_ref, ctx.bias_ref = weight, bias
        ctx.set_materialize_grads(False)
        return y

    @staticmethod
    def backward(ctx, gy):
        a_hi, a_lo, w_hi, w_lo, y = ctx.saved_tensors
        N, K, Np, Kp, M, relu, terms, has_bias, xshape, exact_input = ctx.meta
        if gy is None:
            return (None,) * 6
        gy2 = gy.reshape(M, Np).contiguous()
        want_b = has_bias and ctx.needs_input_grad[2]
        g_hi, g_lo, gb, _ = relu_bwd_split(gy2, y if relu else None, want_b, need_g=False)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # dX = dY W: written straight into [M, K] when the row pitch allows a tensor map (the copy engine drops the
            # columns behind K), else into a padded buffer that is sliced
            direct = K % 4 == 0
            gx = torch.empty((M, K if direct else Kp), dtype=torch.float32, device=gy2.device)
            _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=w_hi, b_lo=w_lo, b_mn=1, alpha=1.0, terms=terms, c=gx, m=M, n=Kp, k=Np, batch=1,
                     inner=1, splits=1, bn=64 if Kp % 128 else 0)
            gx = (gx if direct else gx[:, :K]).reshape(xshape)
        if ctx.needs_input_grad[1]:
            # dW = dY^T X over the rows (reduction padded to 64: rows behind M are TMA zero fill on both operands)
            slot = _grad_slot(ctx.weight_ref)
            kk = _pad_to(M, 64)
            tiles = (_pad_to(Np, 128) // 128) * (Kp // 64)
            splits = _pick_splits(tiles, kk)
            common = dict(a_hi=g_hi, a_lo=g_lo, b_hi=a_hi, b_lo=a_lo, a_mn=1, b_mn=1, alpha=1.0, terms=terms, m=Np, n=Kp, k=kk,
                          batch=1, inner=1, splits=splits, c_row_split=0, c_accumulate=1, bn=64 if Kp % 128 else 0,
                          b_exact=1 if exact_input else 0)
            if slot is not None and K % 4 == 0:
                _gemm_ex(c=slot, **common)               # rows >= N and columns >= K fall outside the tensor map: dropped
            else:
                part = torch.zeros((Np, Kp), dtype=torch.float32, device=gy2.device)
                _gemm_ex(c=part, **common)
                gw = part[:N, :K]
                if slot is not None:
                    slot.add_(gw)
                    gw = None
        if want_b and gb is not None:
            gb = gb[:N]
            bslot = _grad_slot(ctx.bias_ref)
            if bslot is not None:
                bslot.add_(gb)
                gb = None
        return gx, gw, gb, None, None, None


def linear_any(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False, terms: int = 3,
               exact_input: bool = False) -> torch.Tensor:
    """fc_block of any shape on the tensor cores (see _PaddedLinear).  x may still be an integer / fp16 observation tensor;
    exact_input: its values are exactly representable in bf16 (0/1 flags, counts <= 256): one product less, no lo operand."""
    N = weight.shape[0]
    y = _PaddedLinear.apply(x, weight, bias, relu, terms, exact_input)
    return y[:, :N].reshape(*x.shape[:-1], N)


class _GluGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, x):
        gate, x = gate.contiguous(), x.contiguous()
        out = torch.empty_like(x)
        lib.call('dsb_glu_gate_fwd', gate, x, out, x.numel())
        ctx.save_for_backward(gate, x)
        return out

    @staticmethod
    def backward(ctx, go):
        gate, x = ctx.saved_tensors
        dg, dx = torch.empty_like(gate), torch.empty_like(x)
        lib.call('dsb_glu_gate_bwd', go.contiguous(), gate, x, dg, dx, x.numel())
        return dg, dx


def glu_gate(gate: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """sigmoid(gate) * x, the gate of GLU (module_utils.py:508-524), one kernel each way."""
    if _use_kernel(x) and x.numel() % 4 == 0 and gate.shape == x.shape:
        return _GluGate.apply(gate.float(), x.float())
    return torch.sigmoid(gate) * x


class _OneHotLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, idx, relu, embedding, clamp_max, flag):
        if embedding:
            C, N = weight.shape
            so, sc = 1, N
        else:
            N, C = weight.shape
            so, sc = C, 1
        P = idx.numel()
        out = torch.empty((P, N), dtype=torch.float32, device=weight.device)
        lib.call('dsb_onehot_linear_fwd', weight, bias, idx, out, P, N, C, so, sc, 1 if relu else 0, 1 if clamp_max else 0, flag)
        ctx.save_for_backward(idx, out if relu else None)
        ctx.meta = (P, N, C, so, sc, relu, tuple(weight.shape), bias is not None)
        ctx.refs = (weight, bias)
        return out

    @staticmethod
    def backward(ctx, go):
        idx, out = ctx.saved_tensors
        P, N, C, so, sc, relu, wshape, has_bias = ctx.meta
        w_ref, b_ref = ctx.refs
        wslot, bslot = _grad_slot(w_ref), (_grad_slot(b_ref) if has_bias else None)
        gw = wslot if wslot is not None else torch.zeros(wshape, dtype=torch.float32, device=go.device)
        gb = None
        if has_bias:
            gb = bslot if bslot is not None else torch.zeros(N, dtype=torch.float32, device=go.device)
        lib.call('dsb_onehot_linear_bwd', go.contiguous(), out, idx, gw, gb, P, N, C, so, sc, 1 if relu else 0)
        return (None if wslot is not None else gw), (None if (bslot is not None or not has_bias) else gb), None, None, None, None, None


def onehot_linear(weight: torch.Tensor, bias: Optional[torch.Tensor], idx: torch.Tensor, relu: bool = True,
                  embedding: bool = False, clamp_max: bool = False, flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(fc(one_hot(idx))) without the one-hot: a gather of weight columns (fc weight [N, C]) or rows (``embedding``:
    table [C, N]) + bias + ReLU (action_type_head.py:61-63, action_arg_head.py:49-52, scalar_encoder.py:105-116); the
    backward scatters into the weight gradient.  Ids outside [0, C) are clamped; bit 4 of ``flag`` records them unless
    ``clamp_max`` (the scalar encoder clamps too-large ids on purpose, scalar_encoder.py:110-114)."""
    shape = idx.shape
    i64 = idx.reshape(-1).to(torch.int64).contiguous()
    if _use_kernel(weight):
        out = _OneHotLinear.apply(weight, bias, i64, relu, embedding, clamp_max, flag)
    else:
        out = _standin('onehot_linear')(weight, bias, i64, relu, embedding, clamp_max, flag)
    return out.view(*shape, out.shape[-1])


class KeyGradSink:
    """Gradient accumulator of the stacked key projection [P, E, 64] of the two pointer heads.  Its three consumers
    (su_prefix_mean, su_logits, target_unit_logits) each touch only some rows / columns; instead of every one of them
    returning a full zero-padded [P, E, 64] tensor for autograd to add up (three 0.5 GB fills + two adds at P = 4096) their
    backward kernels write into ONE zero-initialised buffer and hand autograd None; ``fork_keys`` returns the buffer as the
    gradient of the projection once all of them have run (autograd runs a node only after every consumer's backward)."""

    def __init__(self):
        self.buf = None

    def get(self, like: torch.Tensor) -> torch.Tensor:
        if self.buf is None:
            self.buf = torch.zeros_like(like)
        return self.buf


class _KeyFork(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kfull, sink):
        ctx.sink = sink
        ctx.set_materialize_grads(False)
        return kfull.view_as(kfull)

    @staticmethod
    def backward(ctx, g):
        buf, ctx.sink.buf = ctx.sink.buf, None
        if buf is None:
            return g, None
        if g is not None:
            buf.add_(g)
        return buf, None


def fork_keys(kfull: torch.Tensor):
    """-> (kfull', sink): pass both to the pointer-head operators."""
    sink = KeyGradSink()
    return _KeyFork.apply(kfull, sink), sink


class _TargetUnit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kfull, col, query, entity_num, temperature, sink):
        P, E, ld = kfull.shape
        kfull, query = kfull.contiguous(), query.contiguous()
        logits = torch.empty((P, E), dtype=torch.float32, device=kfull.device)
        key = kfull.view(-1)[col:]                                  # the head's 32 columns inside the stacked key projection
        lib.call('dsb_target_unit_fwd', key, ld, query, entity_num, logits, P, E, float(temperature))
        ctx.save_for_backward(kfull, query, entity_num)
        ctx.meta = (col, float(temperature))
        ctx.sink = sink
        return logits

    @staticmethod
    def backward(ctx, gl):
        kfull, query, entity_num = ctx.saved_tensors
        col, temperature = ctx.meta
        P, E, ld = kfull.shape
        if ctx.sink is not None:
            gk = ctx.sink.get(kfull)                                # shared buffer: this head owns columns [col, col + 32)
        else:
            gk = torch.zeros_like(kfull) if ld != 32 else torch.empty_like(kfull)
        gq = torch.empty_like(query)
        lib.call('dsb_target_unit_bwd', gl.contiguous(), kfull.view(-1)[col:], ld, query, entity_num, gk.view(-1)[col:], ld, gq,
                 P, E, temperature)
        return (None if ctx.sink is not None else gk), None, gq, None, None, None


def target_unit_logits(kfull: torch.Tensor, col: int, query: torch.Tensor, entity_num: torch.Tensor,
                       temperature: float, sink: Optional[KeyGradSink] = None) -> torch.Tensor:
    """TargetUnitHead (action_arg_head.py:357-361): logits[p, e] = key[p, e] . query[p] with key = kfull[..., col:col+32],
    entities >= entity_num masked to -1e9, divided by the temperature; one warp-level kernel each way."""
    P, E, ld = kfull.shape
    if _use_kernel(kfull) and E % 4 == 0 and ld % 4 == 0 and query.shape[-1] == 32:
        return _TargetUnit.apply(kfull, col, query.float(), entity_num.to(torch.int64).contiguous(), temperature, sink)
    key = kfull[..., col:col + 32]
    logits = torch.matmul(key, query.unsqueeze(-1)).squeeze(-1)
    valid = torch.arange(E, device=kfull.device).unsqueeze(0) < entity_num.unsqueeze(1)
    return logits.masked_fill(~valid, -1e9) / temperature


# ------------------------------------------------------------------------------------------------
# teacher-forced selected-units pointer network (K12 training path, csrc/su_train.cu)
# ------------------------------------------------------------------------------------------------
class _SuPrefixMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kfull, su, entity_num, num, S, sink):
        P, E, ld = kfull.shape
        mean = torch.empty((P, S, 32), dtype=torch.float32, device=kfull.device)
        cnt = torch.empty((P, S), dtype=torch.int32, device=kfull.device)
        lib.call('dsb_su_prefix_mean_fwd', kfull, ld, su, su.shape[1], entity_num, num, mean, cnt, P, E, S)
        ctx.save_for_backward(kfull, su, entity_num, num, cnt)
        ctx.sink, ctx.S = sink, S
        return mean

    @staticmethod
    def backward(ctx, g):
        kfull, su, entity_num, num, cnt = ctx.saved_tensors
        P, E, ld = kfull.shape
        gk = ctx.sink.get(kfull) if ctx.sink is not None else torch.zeros_like(kfull)
        lib.call('dsb_su_prefix_mean_bwd', g.contiguous(), su, su.shape[1], entity_num, num, cnt, gk, ld, P, E, ctx.S)
        return (None if ctx.sink is not None else gk), None, None, None, None, None


class _SuLstm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ig, w_hh, gam_h, bet_h, gam_c, bet_c):
        P, S, _ = ig.shape
        ig = ig.contiguous()
        f32 = dict(dtype=torch.float32, device=ig.device)
        hs, cs = torch.empty((P, S, 32), **f32), torch.empty((P, S, 32), **f32)
        gates, hg = torch.empty((P, S, 128), **f32), torch.empty((P, S, 128), **f32)
        pre_c, st = torch.empty((P, S, 32), **f32), torch.empty((P, S, 4), **f32)
        w = w_hh.detach().contiguous()
        lib.call('dsb_su_lstm_fwd', ig, w, gam_h, bet_h, gam_c, bet_c, hs, cs, gates, hg, pre_c, st, P, S)
        ctx.save_for_backward(w, gam_h, gam_c, bet_c, hs, cs, gates, hg, pre_c, st)
        ctx.refs = (w_hh, gam_h, bet_h, gam_c, bet_c)
        return hs

    @staticmethod
    def backward(ctx, g_hs):
        w, gam_h, gam_c, bet_c, hs, cs, gates, hg, pre_c, st = ctx.saved_tensors
        P, S, _ = hs.shape
        f32 = dict(dtype=torch.float32, device=hs.device)
        d_ig, d_hg = torch.empty((P, S, 128), **f32), torch.empty((P, S, 128), **f32)
        refs = ctx.refs
        slots = [_grad_slot(p) for p in refs[1:]]
        direct = all(s_ is not None for s_ in slots)
        dgh, dbh, dgc, dbc = slots if direct else (torch.zeros(128, **f32), torch.zeros(128, **f32), torch.zeros(32, **f32),
                                                   torch.zeros(32, **f32))
        lib.call('dsb_su_lstm_bwd', g_hs.contiguous(), w, gam_h, gam_c, bet_c, cs, gates, hg, pre_c, st, d_ig, d_hg, dgh, dbh, dgc,
                 dbc, P, S)
        gw = None
        if ctx.needs_input_grad[1]:
            # dW_hh = d_hg^T h_prev over all (row, step) pairs: one (zero-padded) tensor-core GEMM
            h_prev = torch.cat([torch.zeros((P, 1, 32), **f32), hs[:, :-1]], dim=1).reshape(P * S, 32)
            gw = small_weight_grad(d_hg.reshape(P * S, 128), h_prev, accumulate_into=_grad_slot(refs[0]))
        return (d_ig, gw) + ((None,) * 4 if direct else (dgh, dbh, dgc, dbc))


def small_weight_grad(g: torch.Tensor, x: torch.Tensor, accumulate_into: Optional[torch.Tensor] = None):
    """dW [N, K] = g^T x for a skinny layer (N, K small, many rows) on the tcgen05 kernel: both operands packed to zero-padded
    bf16 pairs, reduction split over the rows.  Adds into ``accumulate_into`` ([N, K], K % 4 == 0) when given."""
    M, N = g.shape
    K = x.shape[1]
    Np, Kp = _pad_to(N, 64), _pad_to(K, 64)
    g_hi, g_lo = pack_pair(g, Np)
    x_hi, x_lo = pack_pair(x, Kp)
    kk = _pad_to(M, 64)
    splits = _pick_splits((_pad_to(Np, 128) // 128) * (Kp // 64), kk)
    common = dict(a_hi=g_hi, a_lo=g_lo, b_hi=x_hi, b_lo=x_lo, a_mn=1, b_mn=1, alpha=1.0, terms=3, m=Np, n=Kp, k=kk, batch=1, inner=1,
                  splits=splits, c_row_split=0, c_accumulate=1, bn=64 if Kp % 128 else 0)
    if accumulate_into is not None and K % 4 == 0:
        _gemm_ex(c=accumulate_into, **common)
        return None
    part = torch.zeros((Np, Kp), dtype=torch.float32, device=g.device)
    _gemm_ex(c=part, **common)
    if accumulate_into is not None:
        accumulate_into.add_(part[:N, :K])
        return None
    return part[:N, :K]


class _SuLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hs, kfull, end_emb, su, entity_num, sink):
        P, S, _ = hs.shape
        _, E, ld = kfull.shape
        hs = hs.contiguous()
        end = end_emb.detach().reshape(-1).contiguous()
        logits = torch.empty((P, S, E + 1), dtype=torch.float32, device=hs.device)
        lib.call('dsb_su_logits_fwd', hs, kfull, ld, end, su, su.shape[1], entity_num, logits, P, E, S)
        ctx.save_for_backward(hs, kfull, end, su, entity_num)
        ctx.sink, ctx.end_ref = sink, end_emb
        return logits

    @staticmethod
    def backward(ctx, gl):
        hs, kfull, end, su, entity_num = ctx.saved_tensors
        P, S, _ = hs.shape
        _, E, ld = kfull.shape
        gk = ctx.sink.get(kfull) if ctx.sink is not None else torch.zeros_like(kfull)
        dhs = torch.empty_like(hs)
        slot = _grad_slot(ctx.end_ref)
        d_end = slot.view(-1) if slot is not None else torch.zeros(32, dtype=torch.float32, device=hs.device)
        lib.call('dsb_su_logits_bwd', gl.contiguous(), hs, kfull, ld, end, su, su.shape[1], entity_num, dhs, gk, d_end, P, E, S)
        return dhs, (None if ctx.sink is not None else gk), (None if slot is not None else d_end.view(ctx.end_ref.shape)), None, \
            None, None


def su_prefix_mean(kfull, selected_units, entity_num, selected_units_num, S: int, sink=None):
    """mean / sum of the keys of the units labelled up to each step (action_arg_head.py:196-199) -> [P, S, 32]."""
    return _SuPrefixMean.apply(kfull, selected_units, entity_num, selected_units_num, S, sink)


def su_lstm(ig, w_hh, gam_h, bet_h, gam_c, bet_c):
    """32-wide LayerNorm-LSTM over the S pointer steps from zero state; ig [P, S, 128] = LN_i(q W_ih^T) -> h [P, S, 32]."""
    return _SuLstm.apply(ig, w_hh, gam_h, bet_h, gam_c, bet_c)


def su_logits(hs, kfull, end_embedding, selected_units, entity_num, sink=None):
    """pointer logits [P, S, E + 1] with the reference's mask recurrence (action_arg_head.py:179-195)."""
    return _SuLogits.apply(hs, kfull, end_embedding, selected_units, entity_num, sink)


# ------------------------------------------------------------------------------------------------
# fused (residual +) LayerNorm and the LayerNorm-LSTM cell
# ------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, want_split):
        D = x.shape[-1]
        x2 = x.reshape(-1, D).contiguous()
        rows = x2.shape[0]
        dev = x.device
        r2 = residual.reshape(-1, D).contiguous() if residual is not None else None
        xin = torch.empty_like(x2) if r2 is not None else x2
        y = torch.empty_like(x2)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        hi = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_split else None
        lo = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_split else None
        lib.call('dsb_layernorm_fwd', x2, r2, weight, bias, xin if r2 is not None else None, y, hi, lo, stats, rows, D, 1e-5)
        ctx.save_for_backward(xin, weight, stats)
        ctx.set_materialize_grads(False)
        ctx.weight_ref, ctx.bias_ref = weight, bias
        ctx.has_res = residual is not None
        ctx.shape = x.shape
        y = y.view(x.shape)
        if want_split:
            hi, lo = hi.view(x.shape), lo.view(x.shape)
            ctx.mark_non_differentiable(hi, lo)
            return y, hi, lo
        return y, None, None

    @staticmethod
    def backward(ctx, gy, _ghi, _glo):
        xin, weight, stats = ctx.saved_tensors
        if gy is None:
            return (None,) * 5
        rows, D = xin.shape
        gy2 = gy.reshape(rows, D).contiguous()
        gx = torch.empty_like(xin)
        gw_slot, gb_slot = _grad_slot(ctx.weight_ref), _grad_slot(ctx.bias_ref)
        if gw_slot is not None and gb_slot is not None and ctx.needs_input_grad[2] and ctx.needs_input_grad[3]:
            lib.call('dsb_layernorm_bwd', gy2, xin, weight, stats, gx, gw_slot, gb_slot, 1, rows, D)
            gx = gx.view(ctx.shape)
            return gx, (gx if ctx.has_res else None), None, None, None
        blocks = lib.load().dsb_layernorm_bwd_blocks(rows)
        pg = torch.empty((blocks, D), dtype=torch.float32, device=gy.device)
        pb = torch.empty((blocks, D), dtype=torch.float32, device=gy.device)
        lib.call('dsb_layernorm_bwd', gy2, xin, weight, stats, gx, pg, pb, 0, rows, D)
        gx = gx.view(ctx.shape)
        return gx, (gx if ctx.has_res else None), pg.sum(0), pb.sum(0), None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, residual: Optional[torch.Tensor] = None,
               want_split: bool = False) -> torch.Tensor:
    """nn.LayerNorm(x + residual) in one kernel; with want_split the bf16 (hi, lo) pair of the result is produced in
    the same pass and remembered for the next tensor-core GEMM (ops.attach_split)."""
    D = x.shape[-1]
    if _use_kernel(x) and lib.load().dsb_layernorm_supported(D):
        y, hi, lo = _LayerNorm.apply(x.float(), residual, weight, bias, want_split)
        return attach_split(y, hi, lo) if want_split else y
    if x.is_cuda and residual is None and not want_split and lib.load().dsb_ln_small_supported(D):
        return _LayerNormSmall.apply(x.float(), weight, bias)
    s = x if residual is None else x + residual
    return F.layer_norm(s, (D,), weight, bias, 1e-5)


class _LayerNormSmall(torch.autograd.Function):
    """LayerNorm over narrow rows (D = 32 / 64 / 96): the beginning-build-order transformer (csrc/small_ops.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        D = x.shape[-1]
        x2 = x.reshape(-1, D).contiguous()
        y = torch.empty_like(x2)
        stats = torch.empty((x2.shape[0], 2), dtype=torch.float32, device=x.device)
        lib.call('dsb_ln_small_fwd', x2, weight, bias, y, stats, x2.shape[0], D, 1e-5)
        ctx.save_for_backward(x2, weight, stats)
        ctx.refs, ctx.shape = (weight, bias), x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, weight, stats = ctx.saved_tensors
        rows, D = x2.shape
        gx = torch.empty_like(x2)
        gslot, bslot = _grad_slot(ctx.refs[0]), _grad_slot(ctx.refs[1])
        direct = gslot is not None and bslot is not None
        dg = gslot if direct else torch.zeros(D, dtype=torch.float32, device=x2.device)
        db = bslot if direct else torch.zeros(D, dtype=torch.float32, device=x2.device)
        lib.call('dsb_ln_small_bwd', gy.reshape(rows, D).contiguous(), x2, weight, stats, gx, dg, db, rows, D)
        return gx.view(ctx.shape), (None if direct else dg), (None if direct else db)


class _AttnSmall(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, hd):
        B, S, _ = qkv.shape
        qkv = qkv.contiguous()
        out = torch.empty((B, S, heads * hd), dtype=torch.float32, device=qkv.device)
        lib.call('dsb_attn_small_fwd', qkv, out, B, S, heads, hd)
        ctx.save_for_backward(qkv)
        ctx.meta = (heads, hd)
        return out

    @staticmethod
    def backward(ctx, go):
        (qkv,) = ctx.saved_tensors
        heads, hd = ctx.meta
        B, S, _ = qkv.shape
        g = torch.empty_like(qkv)
        lib.call('dsb_attn_small_bwd', qkv, go.contiguous(), g, B, S, heads, hd)
        return g, None, None


def small_attention(qkv: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """Unmasked multi-head self-attention over short sequences (module_utils.py:88-111 at S = 20, 2 heads of 8: the
    beginning-build-order transformer): qkv [B, S, 3 * heads * hd] -> [B, S, heads * hd], one warp per (sequence, head)."""
    B, S, _ = qkv.shape
    if _use_kernel(qkv) and S <= 32 and hd in (8, 16):
        return _AttnSmall.apply(qkv.float(), heads, hd)
    q, k, v = qkv.view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax(torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd), dim=-1)
    return torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, S, heads * hd)


def bo_tokens(order: torch.Tensor, location: torch.Tensor, spatial_x: int, num_actions: int = 174, Kp: int = 256) -> torch.Tensor:
    """Token features of the beginning-build-order encoder (scalar_encoder.py:33-45) as the exact bf16 operand
    [B * 20, Kp] of its embedding GEMM (every feature is 0 / 1)."""
    B, L = order.shape
    hi = torch.empty((B * L, Kp), dtype=torch.bfloat16, device=order.device)
    lib.call('dsb_bo_tokens', order.to(torch.int16).contiguous(), location.to(torch.int16).contiguous(), spatial_x, hi, B, L,
             num_actions, Kp)
    return hi


class _LstmCell(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ig, hg, c_in, gam_h, bet_h, gam_c, bet_c):
        B, G = ig.shape
        H = G // 4
        dev = ig.device
        ig, hg, c_in = ig.contiguous(), hg.contiguous(), c_in.contiguous()
        h = torch.empty((B, H), dtype=torch.float32, device=dev)
        c = torch.empty_like(h)
        gates = torch.empty((B, G), dtype=torch.float32, device=dev)
        pre_c = torch.empty_like(h)
        st_h = torch.empty((B, 2), dtype=torch.float32, device=dev)
        st_c = torch.empty((B, 2), dtype=torch.float32, device=dev)
        lib.call('dsb_lstm_cell_fwd', ig, hg, c_in, gam_h, bet_h, gam_c, bet_c, h, c, gates, st_h, pre_c, st_c, B, H, 1e-5)
        ctx.save_for_backward(gates, hg, st_h, c_in, pre_c, st_c, gam_h, gam_c, bet_c)
        return h, c

    @staticmethod
    def backward(ctx, gh, gc):
        gates, hg, st_h, c_in, pre_c, st_c, gam_h, gam_c, bet_c = ctx.saved_tensors
        B, G = gates.shape
        H = G // 4
        dev = gates.device
        gh = gh.contiguous() if gh is not None else torch.zeros((B, H), dtype=torch.float32, device=dev)
        gc = gc.contiguous() if gc is not None else None
        d_ig = torch.empty_like(gates)
        d_hg = torch.empty_like(gates)
        d_cin = torch.empty((B, H), dtype=torch.float32, device=dev)
        dgh = torch.zeros(G, dtype=torch.float32, device=dev)
        dbh = torch.zeros(G, dtype=torch.float32, device=dev)
        dgc = torch.zeros(H, dtype=torch.float32, device=dev)
        dbc = torch.zeros(H, dtype=torch.float32, device=dev)
        lib.call('dsb_lstm_cell_bwd', gh, gc, gates, hg, st_h, c_in, pre_c, st_c, gam_h, gam_c, bet_c, d_ig, d_hg, d_cin,
                 dgh, dbh, dgc, dbc, B, H)
        return d_ig, d_hg, d_cin, dgh, dbh, dgc, dbc


class _LstmLayer(torch.autograd.Function):
    """One LayerNorm-LSTM layer over all timesteps (model/lstm.py:120-167) as ONE autograd node and ONE kernel launch each way
    (csrc/lstm_seq.cu): a CTA keeps the h / c of its batch rows on chip for the whole sequence and streams W_hh from L2.
    Outside the kernel remain the two things that are genuinely GEMM shaped: the input projection of all steps (policy_net) and
    the W_hh gradient, one tensor-core GEMM over all L*B rows; LayerNorm parameter gradients are accumulated by the kernel
    straight into the arena."""

    @staticmethod
    def forward(ctx, ig_all, h0, c0, w_hh, gam_h, bet_h, gam_c, bet_c):
        L, B, G = ig_all.shape
        H = G // 4
        dev = ig_all.device
        ig_all, h0, c0 = ig_all.contiguous(), h0.contiguous(), c0.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        hs, cs = torch.empty((L, B, H), **f32), torch.empty((L, B, H), **f32)
        gates, hg_all = torch.empty((L, B, G), **f32), torch.empty((L, B, G), **f32)
        pre_c, st_h, st_c = torch.empty((L, B, H), **f32), torch.empty((L, B, 2), **f32), torch.empty((L, B, 2), **f32)
        w_t = weight_cached(w_hh, 'hh_t', lambda: w_hh.detach().t().contiguous())
        lib.call('dsb_lstm_seq_fwd', ig_all, h0, c0, w_t, gam_h, bet_h, gam_c, bet_c, hs, cs, gates, hg_all, st_h, pre_c, st_c,
                 L, B, H, 1e-5)
        ctx.save_for_backward(gates, hg_all, st_h, pre_c, st_c, hs, cs, h0, c0, w_hh, gam_h, gam_c, bet_c)
        ctx.refs = (w_hh, gam_h, bet_h, gam_c, bet_c)
        ctx.set_materialize_grads(False)
        return hs, cs[L - 1].clone()

    @staticmethod
    def backward(ctx, g_hs, g_clast):
        gates, hg_all, st_h, pre_c, st_c, hs, cs, h0, c0, w_hh, gam_h, gam_c, bet_c = ctx.saved_tensors
        L, B, G = gates.shape
        H = G // 4
        dev = gates.device
        f32 = dict(dtype=torch.float32, device=dev)
        d_ig, d_hg = torch.empty((L, B, G), **f32), torch.empty((L, B, G), **f32)
        dh0, dc0 = torch.empty((B, H), **f32), torch.empty((B, H), **f32)
        refs = ctx.refs
        slots = [_grad_slot(p) for p in refs[1:]]
        direct = all(s_ is not None for s_ in slots)
        if direct:
            dgh, dbh, dgc, dbc = slots
        else:
            dgh, dbh = torch.zeros(G, **f32), torch.zeros(G, **f32)
            dgc, dbc = torch.zeros(H, **f32), torch.zeros(H, **f32)
        lib.call('dsb_lstm_seq_bwd', g_hs.contiguous() if g_hs is not None else None,
                 g_clast.contiguous() if g_clast is not None else None, gates, hg_all, st_h, pre_c, st_c, cs, c0, w_hh.detach(),
                 gam_h, gam_c, bet_c, d_ig, d_hg, dh0, dc0, dgh, dbh, dgc, dbc, L, B, H)
        gw = None
        if ctx.needs_input_grad[3]:
            h_prev = torch.cat([h0.unsqueeze(0), hs[:L - 1]], dim=0).reshape(L * B, H)
            dg2 = d_hg.reshape(L * B, G)
            if (L * B) % 64 == 0 and L * B >= 128 and G % 128 == 0 and H % 128 == 0:
                g_hi, g_lo = split_bf16(dg2)
                x_hi, x_lo = split_bf16(h_prev)
                gw = weight_grad(g_hi, g_lo, x_hi, x_lo, 3, accumulate_into=_grad_slot(refs[0]))
            else:
                gw = dg2.t() @ h_prev
        return (d_ig, dh0, dc0, gw) + ((None,) * 4 if direct else (dgh, dbh, dgc, dbc))


def lstm_layer(ig_all, h0, c0, w_hh, gam_h, bet_h, gam_c, bet_c):
    """LayerNorm-LSTM layer over [L,B,4H] pre-computed input gates -> (h for every step [L,B,H], c after the last step)."""
    return _LstmLayer.apply(ig_all, h0, c0, w_hh, gam_h, bet_h, gam_c, bet_c)


def lstm_cell(ig, hg, c_in, gam_h, bet_h, gam_c, bet_c):
    """LayerNormLSTMCell (model/lstm.py:138-153) after the two matmuls, fused: returns (h', c')."""
    H = c_in.shape[-1]
    if _use_kernel(ig) and H in (128, 384):
        return _LstmCell.apply(ig, hg, c_in, gam_h, bet_h, gam_c, bet_c)
    G = 4 * H
    gates = ig + F.layer_norm(hg, (G,), gam_h, bet_h, 1e-5)
    i, f, g, o = gates.chunk(4, 1)
    c2 = F.layer_norm(torch.sigmoid(f) * c_in + torch.sigmoid(i) * torch.tanh(g), (H,), gam_c, bet_c, 1e-5)
    return torch.sigmoid(o) * torch.tanh(c2), c2


# ------------------------------------------------------------------------------------------------
# implicit-GEMM convolutions on NHWC activations (spatial ResNet K7, location head K14)
# ------------------------------------------------------------------------------------------------
def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _conv_weight_matrix(weight: torch.Tensor, cin_pad: int, cout_pad: int) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [cout_pad, kh*kw*cin_pad] with the reduction index ordered (ky, kx, cin)."""
    Cout, Cin, kh, kw = weight.shape
    w = weight.permute(0, 2, 3, 1)
    w = F.pad(w, (0, cin_pad - Cin, 0, 0, 0, 0, 0, cout_pad - Cout))
    return w.reshape(cout_pad, kh * kw * cin_pad).contiguous()


def _conv_forward_operands(weight, cin_pad, cout_pad):
    wm = _conv_weight_matrix(weight.detach(), cin_pad, cout_pad)
    return (wm,) + tuple(split_bf16(wm))


def _conv_dx_operands(wm, cout_pad, kh, kw, C):
    """W'[cin, (ky,kx,cout)] = W[cout, cin, kh-1-ky, kw-1-kx] as a bf16 pair (the input-gradient convolution's weights)"""
    w4 = wm.view(cout_pad, kh, kw, C).flip(1, 2).permute(3, 1, 2, 0).reshape(C, kh * kw * cout_pad).contiguous()
    return split_bf16(w4)


class _ConvNHWC(torch.autograd.Function):
    """y = act(conv(x, w) + b [+ residual]) for 3x3 (pad 1) and 1x1 kernels, NHWC, channels padded to 64.

    Forward, input gradient (same kernel with flipped / transposed weights) and weight gradient (activation read
    tap-shifted as an MN-major operand, split-K over pixels) all run on the tcgen05 GEMM: the im2col matrix is never
    materialised, halos come from TMA zero fill."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu, terms, emit_split=False):
        N, H, W, C = x.shape
        Cout, Cin, kh, kw = weight.shape
        taps = kh * kw
        cout_pad = _pad_to(Cout, 64)
        wm, w_hi, w_lo = weight_cached(weight, 'conv_fwd_%d_%d' % (C, cout_pad), lambda: _conv_forward_operands(weight, C, cout_pad))
        x_hi, x_lo = split_bf16(x)
        b = F.pad(bias, (0, cout_pad - Cout)).contiguous() if bias is not None else None
        pair_only = emit_split == 'only'
        y = torch.empty((N * H * W, cout_pad), dtype=torch.float32, device=x.device) if not pair_only else None
        y_hi = torch.empty((N * H * W, cout_pad), dtype=torch.bfloat16, device=x.device) if emit_split else None
        y_lo = torch.empty((N * H * W, cout_pad), dtype=torch.bfloat16, device=x.device) if emit_split else None
        res = residual.reshape(N * H * W, cout_pad).contiguous() if residual is not None else None
        _gemm_ex(a_hi=x_hi, a_lo=x_lo, b_hi=w_hi, b_lo=w_lo, bias=b, residual=res, alpha=1.0, relu=1 if relu else 0,
                 terms=terms, c=y, c_hi=y_hi, c_lo=y_lo, m=N * H * W, n=cout_pad, k=taps * C, batch=1, inner=1, splits=1,
                 a_conv=1, conv_h=H, conv_w=W, conv_c=C, conv_taps=taps, conv_imgs=N)
        if pair_only:
            y = pair_only_placeholder((N * H * W, cout_pad), x.device)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x_hi, x_lo, wm, (y_hi if emit_split else y) if relu else None)
        ctx.meta = (N, H, W, C, Cout, Cin, kh, kw, cout_pad, relu, terms, bias is not None, residual is not None)
        ctx.bias_ref, ctx.weight_ref = bias, weight
        if emit_split:
            y_hi, y_lo = y_hi.view(N, H, W, cout_pad), y_lo.view(N, H, W, cout_pad)
            ctx.mark_non_differentiable(y_hi, y_lo)
            return y.view(N, H, W, cout_pad), y_hi, y_lo
        return y.view(N, H, W, cout_pad), None, None

    @staticmethod
    def backward(ctx, gy, _ghi=None, _glo=None):
        x_hi, x_lo, wm, y = ctx.saved_tensors
        N, H, W, C, Cout, Cin, kh, kw, cout_pad, relu, terms, has_bias, has_res = ctx.meta
        if gy is None:
            return (None,) * 7
        taps = kh * kw
        gy2 = gy.reshape(N * H * W, cout_pad).contiguous()
        want_b = has_bias and ctx.needs_input_grad[2]
        b_slot = _grad_slot(ctx.bias_ref) if (want_b and Cout == cout_pad) else None
        g_hi, g_lo, gb_full, g = relu_bwd_split(gy2, y if relu else None, want_b, need_g=has_res and relu,
                                                bias_grad_out=b_slot)
        gx = gw = gb = gres = None
        if has_res:
            gres = (g if relu else gy2).view(N, H, W, cout_pad)
        if ctx.needs_input_grad[0]:
            # dX = conv(dY, W') with W'[cin, (ky,kx,cout)] = W[cout, cin, kh-1-ky, kw-1-kx]
            wt_hi, wt_lo = weight_cached(ctx.weight_ref, 'conv_dx_%d_%d' % (C, cout_pad),
                                         lambda: _conv_dx_operands(wm, cout_pad, kh, kw, C))
            gx = torch.empty((N * H * W, C), dtype=torch.float32, device=gy2.device)
            _gemm_ex(a_hi=g_hi.view(N, H, W, cout_pad), a_lo=g_lo.view(N, H, W, cout_pad), b_hi=wt_hi, b_lo=wt_lo,
                     alpha=1.0, terms=terms, c=gx, m=N * H * W, n=C, k=taps * cout_pad, batch=1, inner=1, splits=1,
                     a_conv=1, conv_h=H, conv_w=W, conv_c=cout_pad, conv_taps=taps, conv_imgs=N)
            gx = gx.view(N, H, W, C)
        if ctx.needs_input_grad[1]:
            pix = N * H * W
            m_pad = _pad_to(cout_pad, 128)
            n = taps * C
            bn = 128 if n % 128 == 0 else 64
            splits = _pick_splits((m_pad // 128) * (n // bn), pix)
            part = torch.zeros((m_pad, n), dtype=torch.float32, device=gy2.device)
            _gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=x_hi, b_lo=x_lo, a_mn=1, b_mn=1, b_conv=1, alpha=1.0, terms=terms, c=part,
                     m=cout_pad, n=n, k=pix, batch=1, inner=1, splits=splits, c_row_split=0, c_accumulate=1, bn=bn,
                     conv_h=H, conv_w=W, conv_c=C, conv_taps=taps, conv_imgs=N)
            gwm = part[:Cout]
            gw = gwm.view(Cout, kh, kw, C)[..., :Cin].permute(0, 3, 1, 2).contiguous()
        if want_b and gb_full is not None:
            gb = gb_full[:Cout]
        return gx, gw, gb, gres, None, None, None


def conv_geometry_supported(H: int, W: int, C: int) -> bool:
    """Map sizes the implicit-GEMM convolution covers: its TMA boxes are whole image rows of a power-of-two width <= 64
    (gemm_tcgen05.cu make_map_nhwc).  Other sizes (the reference default 160x152 gives 76x80 / 38x40 / 19x20 maps) take the
    fp32 library convolution for 3x3 kernels and the fc path for 1x1 kernels."""
    return C % 64 == 0 and W <= 64 and 64 % W == 0 and H % max(1, 128 // W) == 0


def _conv1x1_as_fc(x, weight, bias, relu, residual, terms):
    """A 1x1 convolution at a map size the implicit GEMM's TMA boxes do not cover (e.g. 19x20 at the 160x152 default): it is a
    plain fc over the pixels, so it stays on the tensor cores; the residual / ReLU are applied after it."""
    N, H, W, C = x.shape
    Cout, Cin = weight.shape[:2]
    cp = _pad_to(Cout, 64)
    w2 = F.pad(weight.reshape(Cout, Cin), (0, C - Cin, 0, cp - Cout))
    b2 = F.pad(bias, (0, cp - Cout)) if bias is not None else None
    x2 = x.reshape(N * H * W, C)
    sp = getattr(x, '_dsb_split', None)
    if sp is not None and sp[0].shape == x.shape:
        x2 = attach_split(x2, sp[0].reshape(x2.shape), sp[1].reshape(x2.shape))
    y = linear(x2, w2, b2, relu and residual is None, terms, allow_n64=True).view(N, H, W, cp)
    if residual is not None:
        y = y + residual
        y = torch.relu(y) if relu else y
    return y


def conv_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
              residual: Optional[torch.Tensor] = None, terms: int = 3, emit_split: bool = False) -> torch.Tensor:
    """conv2d_block (ctools/torch_utils/network/nn_module.py:119-174) on an NHWC activation [N,H,W,C] whose channel
    count is a multiple of 64 (zero padded); weight is the reference's [Cout, Cin<=C, k, k] (k = 1 or 3, padding k//2).
    Returns [N,H,W,pad64(Cout)] (padded output channels are exactly 0)."""
    N, H, W, C = x.shape
    kh = weight.shape[2]
    if _use_kernel(x) and conv_geometry_supported(H, W, C):
        if getattr(x, '_dsb_split', None) is None:
            x = x.contiguous()
        y, y_hi, y_lo = _ConvNHWC.apply(x, weight, bias, residual, relu, terms, emit_split)
        return attach_split(y, y_hi, y_lo) if emit_split else y
    if _use_kernel(x) and kh == 1 and C % 64 == 0:
        return _conv1x1_as_fc(x, weight, bias, relu, residual, terms)
    Cout, Cin = weight.shape[:2]
    y = F.conv2d(x[..., :Cin].permute(0, 3, 1, 2), weight, bias, padding=kh // 2).permute(0, 2, 3, 1)
    y = F.pad(y, (0, _pad_to(Cout, 64) - Cout))
    if residual is not None:
        y = y + residual
    y = torch.relu(y) if relu else y
    return y.contiguous()          # channels-last in memory, like the kernel path: downstream kernels take raw pointers


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


class _Upsample2xNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, H, W, C = x.shape
        x = x.contiguous()
        out = torch.empty((N, 2 * H, 2 * W, C), dtype=torch.float32, device=x.device)
        lib.call('dsb_upsample_bilinear2x_nhwc_fwd', x, out, N, H, W, C)
        ctx.shape = (N, H, W, C)
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = ctx.shape
        gin = torch.empty((N, H, W, C), dtype=torch.float32, device=g.device)
        lib.call('dsb_upsample_bilinear2x_nhwc_bwd', g.contiguous(), gin, N, H, W, C)
        return gin


def upsample_bilinear2x_nhwc(x: torch.Tensor) -> torch.Tensor:
    """bilinear x2 (align_corners=False) on a channels-last activation [N,H,W,C]."""
    if _use_kernel(x):
        return _Upsample2xNHWC.apply(x.float())
    return _standin('upsample_bilinear2x_nhwc')(x)


class _UpShift9(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias):
        N, H, W, _ = z.shape
        z = z.contiguous()
        out = torch.empty((N, 2 * H, 2 * W), dtype=torch.float32, device=z.device)
        lib.call('dsb_upshift9_fwd', z, bias, out, N, H, W)
        ctx.shape = (N, H, W)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W = ctx.shape
        g = g.contiguous()
        gz = torch.empty((N, H, W, 9), dtype=torch.float32, device=g.device)
        lib.call('dsb_upshift9_bwd', g, gz, N, H, W)
        return gz, (g.sum().reshape(1) if ctx.has_bias else None)


class _Proj9(torch.autograd.Function):
    """z = x @ wm for x [..., 32], wm [32, 9] (csrc/upsample.cu: dsb_proj9_*): the library GEMMs are very slow at this shape."""

    @staticmethod
    def forward(ctx, x, wm):
        x = x.contiguous()
        wm = wm.contiguous()
        pixels = x.numel() // 32
        z = torch.empty((*x.shape[:-1], 9), dtype=torch.float32, device=x.device)
        lib.call('dsb_proj9_fwd', x, wm, z, pixels, 32)
        ctx.save_for_backward(x, wm)
        return z

    @staticmethod
    def backward(ctx, gz):
        x, wm = ctx.saved_tensors
        gx = torch.empty_like(x)
        gw = torch.zeros((32, 9), dtype=torch.float32, device=x.device)
        lib.call('dsb_proj9_bwd', x, gz.contiguous(), wm, gx, gw, x.numel() // 32, 32)
        return gx, gw


def upsample_conv3x3_single(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """conv2d(F.interpolate(x, 2, 'bilinear'), weight[1,C,3,3], bias, padding=1) for a channels-last x [N,H,W,Cpad>=C]
    -> [N, 2H, 2W].  Up-sampling commutes with the channel contraction, so the 3x3 taps are applied as nine shifted
    up-samplings of a 9-channel low-resolution projection z = x . w (see csrc/upsample.cu)."""
    C = weight.shape[1]
    wm = weight[0].reshape(C, 9)
    if x.shape[-1] > C:
        wm = F.pad(wm, (0, 0, 0, x.shape[-1] - C))
    if x.is_cuda and x.shape[-1] == 32 and not _HOST_LOGIC_TESTING:
        z = _Proj9.apply(x, wm)                               # [N,H,W,9]
    else:
        z = torch.matmul(x, wm)                               # [N,H,W,9]  (small library GEMM, N = 9)
    if _use_kernel(x):
        return _UpShift9.apply(z, bias)
    return _standin('upshift9')(z, bias)


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, want_split):
        N, H, W, C = x.shape
        x = x.contiguous()
        oshape = (N, H // 2, W // 2, C)
        out = torch.empty(oshape, dtype=torch.float32, device=x.device)
        hi = torch.empty(oshape, dtype=torch.bfloat16, device=x.device) if want_split else None
        lo = torch.empty(oshape, dtype=torch.bfloat16, device=x.device) if want_split else None
        idx = torch.empty(oshape, dtype=torch.uint8, device=x.device)
        lib.call('dsb_maxpool2_nhwc_fwd', x, out, hi, lo, idx, N, H, W, C)
        ctx.save_for_backward(idx)
        ctx.xshape = x.shape
        ctx.set_materialize_grads(False)
        if want_split:
            ctx.mark_non_differentiable(hi, lo)
        return out, hi, lo

    @staticmethod
    def backward(ctx, gout, _ghi=None, _glo=None):
        (idx,) = ctx.saved_tensors
        if gout is None:
            return None, None
        N, H, W, C = ctx.xshape
        gx = torch.empty(ctx.xshape, dtype=torch.float32, device=gout.device)
        lib.call('dsb_maxpool2_nhwc_bwd', gout.contiguous(), idx, gx, N, H, W, C)
        return gx, None


def max_pool2_nhwc(x: torch.Tensor, want_split: bool = True) -> torch.Tensor:
    """F.max_pool2d(x, 2, 2) on a channels-last [N,H,W,C] activation; the bf16 pair of the result is attached."""
    if _use_kernel(x) and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0 and x.shape[3] % 4 == 0:
        out, hi, lo = _MaxPool2.apply(x, want_split)
        return attach_split(out, hi, lo) if want_split else out
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


class _GateUpdate(torch.autograd.Function):
    """relu(tanh(r * sigmoid(g)) * sp + x) [+ skip] in one kernel each way (csrc/optim.cu: dsb_gate_update_*)."""

    @staticmethod
    def forward(ctx, r, g, x, sp, skip, want_split):
        r, g, x = r.contiguous(), g.contiguous(), x.contiguous()
        out = torch.empty_like(x)
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_split else None
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_split else None
        lib.call('dsb_gate_update_fwd', r, g, x, skip.contiguous() if skip is not None else None, sp, out, hi, lo, x.numel())
        ctx.save_for_backward(r, g, x, sp)
        ctx.sp_ref, ctx.has_skip = sp, skip is not None
        ctx.set_materialize_grads(False)
        if want_split:
            ctx.mark_non_differentiable(hi, lo)
        return out, hi, lo

    @staticmethod
    def backward(ctx, gout, _ghi=None, _glo=None):
        r, g, x, sp = ctx.saved_tensors
        if gout is None:
            return (None,) * 6
        gout = gout.contiguous()
        dr, dg, dx = torch.empty_like(r), torch.empty_like(g), torch.empty_like(x)
        slot = _grad_slot(ctx.sp_ref)
        dsp = slot if slot is not None else torch.zeros(1, dtype=torch.float32, device=x.device)
        lib.call('dsb_gate_update_bwd', gout, r, g, x, sp, dr, dg, dx, dsp, x.numel())
        return dr, dg, dx, (None if slot is not None else dsp), (gout if ctx.has_skip else None), None


def gate_update(r: torch.Tensor, g: torch.Tensor, x: torch.Tensor, sp: torch.Tensor, skip: Optional[torch.Tensor] = None,
                want_split: bool = True) -> torch.Tensor:
    """GatedResBlock tail (module_utils.py:228-229): relu(tanh(r * sigmoid(g)) * sp + x), plus ``skip`` (the `x +
    map_skip` that opens the next block of the location head) when given.  With want_split the bf16 pair of the result
    is attached for the convolutions that read it."""
    if _use_kernel(x) and x.numel() % 4 == 0:
        out, hi, lo = _GateUpdate.apply(r, g, x, sp, skip, want_split)
        return attach_split(out, hi, lo) if want_split else out
    out = torch.relu(torch.tanh(r * torch.sigmoid(g)) * sp + x)
    return out + skip if skip is not None else out


class _UpConv(torch.autograd.Function):
    """y = act(conv3x3(upsample_bilinear2x(x), w) + b), channels-last, through the low-resolution factorisation
    z = x . w (tensor-core GEMM, N = 9*C) followed by nine shifted up-samplings (csrc/upsample.cu: dsb_upconv_*)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, pair_only, x_hi=None, x_lo=None):
        N, H, W, Cp = x.shape
        C, Cin = weight.shape[:2]
        M = N * H * W
        ldz = _pad_to(9 * C, 64)
        dev = x.device
        def build():
            wz = torch.zeros((_pad_to(ldz, 128), Cp), dtype=torch.float32, device=dev)
            wz[:9 * C, :Cin] = weight.detach().permute(2, 3, 0, 1).reshape(9 * C, Cin)          # row = tap * C + co
            return split_bf16(wz)
        if x_hi is None:
            x_hi, x_lo = split_bf16(x.reshape(M, Cp).contiguous())
        x_hi, x_lo = x_hi.reshape(M, Cp), x_lo.reshape(M, Cp)
        w_hi, w_lo = weight_cached(weight, 'upconv_%d' % Cp, build)
        z = torch.empty((M, ldz), dtype=torch.float32, device=dev)
        _gemm_ex(a_hi=x_hi, a_lo=x_lo, b_hi=w_hi, b_lo=w_lo, alpha=1.0, terms=3, c=z, m=M, n=ldz, k=Cp, batch=1, inner=1,
                 splits=1, bn=64)
        oshape = (N, 2 * H, 2 * W, C)
        y = None if pair_only else torch.empty(oshape, dtype=torch.float32, device=dev)
        y_hi = torch.empty(oshape, dtype=torch.bfloat16, device=dev) if pair_only else None
        y_lo = torch.empty(oshape, dtype=torch.bfloat16, device=dev) if pair_only else None
        lib.call('dsb_upconv_fwd', z, ldz, bias, 1 if relu else 0, y, y_hi, y_lo, C, N, H, W, C)
        del z
        ctx.save_for_backward(x_hi, x_lo, w_hi, w_lo, (y_hi if pair_only else y) if relu else None)
        ctx.meta = (N, H, W, Cp, C, Cin, ldz, relu, bias is not None)
        ctx.set_materialize_grads(False)
        if pair_only:
            ctx.mark_non_differentiable(y_hi, y_lo)
            return pair_only_placeholder(oshape, dev), y_hi, y_lo
        return y, None, None

    @staticmethod
    def backward(ctx, gy, _ghi=None, _glo=None):
        x_hi, x_lo, w_hi, w_lo, mask = ctx.saved_tensors
        N, H, W, Cp, C, Cin, ldz, relu, has_bias = ctx.meta
        if gy is None:
            return (None,) * 7
        M = N * H * W
        dev = gy.device
        gy2 = gy.reshape(4 * M, C).contiguous()
        want_b = has_bias and ctx.needs_input_grad[2]
        if relu or want_b:
            _, _, gb, g = relu_bwd_split(gy2, mask.reshape(4 * M, C) if relu else None, want_b, need_g=relu, need_split=False)
            g = g if relu else gy2
        else:
            g, gb = gy2, None
        gz_hi = torch.empty((M, ldz), dtype=torch.bfloat16, device=dev)
        gz_lo = torch.empty((M, ldz), dtype=torch.bfloat16, device=dev)
        lib.call('dsb_upconv_bwd', g, C, gz_hi, gz_lo, ldz, N, H, W, C)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty((M, Cp), dtype=torch.float32, device=dev)
            _gemm_ex(a_hi=gz_hi, a_lo=gz_lo, b_hi=w_hi, b_lo=w_lo, b_mn=1, alpha=1.0, terms=3, c=gx, m=M, n=Cp, k=ldz,
                     batch=1, inner=1, splits=1, bn=64 if Cp % 128 else 128)
            gx = gx.view(N, H, W, Cp)
        if ctx.needs_input_grad[1]:
            rows = w_hi.shape[0]                                    # ldz padded to a whole 128-row tile
            bn = 64 if Cp % 128 else 128
            splits = _pick_splits((rows // 128) * (Cp // bn), M)
            gwz = torch.zeros((rows, Cp), dtype=torch.float32, device=dev)
            _gemm_ex(a_hi=gz_hi, a_lo=gz_lo, b_hi=x_hi, b_lo=x_lo, a_mn=1, b_mn=1, alpha=1.0, terms=3, c=gwz, m=ldz, n=Cp, k=M,
                     batch=1, inner=1, splits=splits, c_row_split=0, c_accumulate=1, bn=bn)
            gw = gwz[:9 * C, :Cin].reshape(3, 3, C, Cin).permute(2, 3, 0, 1).contiguous()
        return gx, gw, gb, None, None, None, None


def upsample_conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True,
                     pair_only: bool = False) -> torch.Tensor:
    """conv2d_block(F.interpolate(x, 2, 'bilinear'), weight[C,Cin,3,3], bias, padding=1, activation) on a channels-last x
    [N,H,W,Cpad>=Cin] -> [N,2H,2W,C] (exactly C channels; C in {32, 64} on the GPU).  With pair_only the result exists only
    as the bf16 (hi, lo) pair the next tensor-core GEMM reads (see pair_only_placeholder)."""
    C, Cin = weight.shape[:2]
    if _use_kernel(x) and C in (32, 64) and x.shape[-1] % 64 == 0 and x.shape[1] % 8 == 0 and x.shape[2] % 8 == 0:
        sp = getattr(x, '_dsb_split', None)
        if sp is None or sp[0].shape != x.shape:
            sp = (None, None)
        y, y_hi, y_lo = _UpConv.apply(x, weight, bias, relu, pair_only, sp[0], sp[1])
        return attach_split(y, y_hi, y_lo) if pair_only else y
    up = F.interpolate(x[..., :Cin].permute(0, 3, 1, 2), scale_factor=2., mode='bilinear')
    y = F.conv2d(up, weight, bias, padding=1).permute(0, 2, 3, 1)
    return (torch.relu(y) if relu else y).contiguous()


def upsample_bilinear2x(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(x, scale_factor=2., mode='bilinear') (align_corners=False) on [N,C,H,W] fp32."""
    if _use_kernel(x):
        return _Upsample2x.apply(x.float())
    return _standin('upsample_bilinear2x')(x)


# ------------------------------------------------------------------------------------------------
# flat-arena optimiser (K20)
# ------------------------------------------------------------------------------------------------
class FlatAdam(torch.optim.Optimizer):
    """Gradient clip + Adam over one contiguous fp32 arena: two kernels per step (norm, clip+Adam).

    Reference: RLLearner._setup_optimizer rl_learner.py:73-80 (Adam(betas=(0, 0.99), eps=1e-5)), BaseLearner._setup_optimizer
    base_learner.py:157-181 (SL: Adam(lr, weight_decay) + warm-up / MultiStepLR schedulers), GradClip
    ctools/torch_utils/grad_clip.py:20-144 applied in rl_learner.py:125 / sl_learner.py:70.
    ``grad_scale`` folds DistModule.sync_gradients' division by world size (dist_helper.py:421-431).

    It IS a ``torch.optim.Optimizer`` (one param group holding the arena) so the reference's own plumbing works on it:
    ``MultiStepLR(optimizer)`` / ``GradualWarmupScheduler`` change ``param_groups[0]['lr']`` which ``step`` reads,
    ``update_config`` sets ``g['lr']`` (rl_learner.py:205-206), and the checkpoint hooks call ``state_dict()`` /
    ``load_state_dict()`` (checkpoint_helper.py:124-131,254).  With ``layout`` (Model.optimizer_layout()) the state dict is in
    the reference's per-parameter format ({'state': {param index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]},
    indices = position in model.parameters()), so optimizer states interchange with reference checkpoints in both directions.

    clip_type (GradClip): 'pytorch_norm' / 'clip_norm' = clip_grad_norm_(max_norm); 'none' = norm only; 'momentum_norm' = what
    the reference's implementation DOES: its per-tensor norm momenta are appended to the end of the list instead of stored at
    their index (grad_clip.py:100-103), so ``norm_mom[idx]`` stays None, no gradient is ever scaled and apply() returns the
    global 2-norm - i.e. 'none'.  ``owner``: the Model whose arenas these are; if it is moved (.cuda() / .to()) after the
    optimiser was built the new arenas are picked up (and the moments moved) at the next step.
    """

    def __init__(self, param: torch.Tensor, grad: torch.Tensor, lr: float, betas=(0.0, 0.99), eps: float = 1e-5,
                 max_norm: Optional[float] = 1.0, weight_decay: float = 0.0, clip_type: str = 'pytorch_norm',
                 layout=None, owner=None):
        assert param.is_contiguous() and grad.is_contiguous() and param.numel() == grad.numel()
        assert clip_type in ('pytorch_norm', 'clip_norm', 'none', 'momentum_norm'), clip_type
        self.param, self.grad = param, grad
        self.clip_type = clip_type
        self.max_norm = max_norm if clip_type in ('pytorch_norm', 'clip_norm') else None
        self.layout, self.owner = layout, owner
        super().__init__([param], dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))
        self.exp_avg = torch.zeros_like(param)
        self.exp_avg_sq = torch.zeros_like(param)
        self.t = 0
        self.norm = torch.zeros(1, dtype=torch.float32, device=param.device)
        self._partial = None

    # ---- attribute view of the single param group (bench / tests read and set these)
    @property
    def lr(self):
        return self.param_groups[0]['lr']

    @lr.setter
    def lr(self, v):
        self.param_groups[0]['lr'] = v

    @property
    def betas(self):
        return self.param_groups[0]['betas']

    @property
    def eps(self):
        return self.param_groups[0]['eps']

    def _rebind(self):
        """The owner's arenas were replaced (Model._apply): follow them and carry the moments along."""
        new_p, new_g = self.owner.flat_param, self.owner.flat_grad
        if new_p is self.param and new_g is self.grad:
            return
        assert new_p.numel() == self.param.numel()
        self.param, self.grad = new_p, new_g
        self.param_groups[0]['params'] = [new_p]
        self.exp_avg, self.exp_avg_sq = self.exp_avg.to(new_p.device), self.exp_avg_sq.to(new_p.device)
        self.norm = self.norm.to(new_p.device)
        self._partial = None

    def zero_grad(self, set_to_none: bool = False):
        if self.owner is not None:
            self._rebind()
        self.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, skip_flag: Optional[torch.Tensor] = None):
        """skip_flag: device float[>=1]; non-zero = leave the weights and moments untouched (see dsb_adam_step)."""
        if self.owner is not None:
            self._rebind()
        self.t += 1
        WEIGHT_EPOCH[0] += 1            # the arena is about to change under every cached weight form
        n = self.param.numel()
        g0 = self.param_groups[0]
        lr, (b1, b2), eps, wd = float(g0['lr']), g0['betas'], float(g0['eps']), float(g0.get('weight_decay', 0.0))
        if _use_kernel(self.param):
            if self._partial is None:
                self._partial = torch.empty(lib.load().dsb_sumsq_partials(), dtype=torch.float32, device=self.param.device)
            lib.call('dsb_grad_norm', self.grad, n, self._partial, self.norm, float(grad_scale))
            lib.call('dsb_adam_step', self.param, self.grad, self.exp_avg, self.exp_avg_sq, n, self.norm,
                     float(self.max_norm or 0.0), float(grad_scale), lr, float(b1), float(b2), eps, wd, int(self.t), None, None,
                     skip_flag)
            return self.norm
        return _standin('flat_adam_step')(self, grad_scale, skip_flag, lr, b1, b2, eps, wd)

    # ---- checkpoint interchange (checkpoint_helper.py:85-140,254: {'model', 'optimizer', 'last_iter'})
    def state_dict(self):
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        if self.layout is None:
            return {'flat': True, 'step': self.t, 'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(),
                    'param_groups': [dict(group, params=[0])]}
        state = {}
        if self.t > 0:
            for idx, off, n, shape in self.layout['slots']:
                state[idx] = {'step': torch.tensor(float(self.t)), 'exp_avg': self.exp_avg[off:off + n].view(shape).clone(),
                              'exp_avg_sq': self.exp_avg_sq[off:off + n].view(shape).clone()}
        return {'state': state, 'param_groups': [dict(group, params=list(range(self.layout['num_params'])))]}

    def load_state_dict(self, sd):
        group = sd['param_groups'][0]
        for k in ('lr', 'betas', 'eps', 'weight_decay'):
            if k in group:
                self.param_groups[0][k] = tuple(group[k]) if k == 'betas' else group[k]
        if sd.get('flat'):
            self.t = int(sd['step'])
            self.exp_avg.copy_(sd['exp_avg'])
            self.exp_avg_sq.copy_(sd['exp_avg_sq'])
            return
        assert self.layout is not None, 'a per-parameter optimizer state needs the model layout (Model.optimizer_layout())'
        state = sd['state']
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for idx, off, n, shape in self.layout['slots']:
            st = state.get(idx, state.get(str(idx)))
            if st is None:
                continue
            steps.add(int(float(st['step'])))
            self.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
        # one bias-correction step for the whole arena: the reference steps every parameter together, so they agree
        assert len(steps) <= 1, 'per-parameter Adam steps differ: %s' % sorted(steps)
        self.t = steps.pop() if steps else 0
