"""CPU oracle: a functional fp32 restatement of DI-star's AlphaStar policy hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; the product package
`distar_b200` never does (and fails loudly when its CUDA library is missing).

What it is: plain-PyTorch (CPU, fp32) functions over a flat ``{state_dict key: tensor}`` mapping that
restate, step by step, what the reference computes on the path SURVEY.md §8(a) lists.  The parameter
names are the reference's own ``state_dict`` keys so the same weights drive the reference, this
oracle and the CUDA product.  Each function cites the reference file:line it follows
(paths relative to ``distar/agent/default/`` of opendilab/DI-star @ 12b1c69).

Pinning: the reference ships no golden vectors or KATs for this path, so the oracle
is pinned against *the reference itself executed in the authoring container*
(``tests/test_oracle_vs_reference.py``, skipped where /root/reference is absent) and against the
committed fixtures in ``tests/golden/`` that ``oracle/make_golden.py`` dumped from that run.

``set_matmul_emulation`` optionally rounds matmul/conv operands (tf32 / bf16 / bf16x3) so precision
modes of the CUDA path can be studied on CPU; it is off ('fp32') for every parity use.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# --------------------------------------------------------------------------------------------
# Static tables restated from model/actor_critic_default_config.yaml and lib/features.py
# --------------------------------------------------------------------------------------------
MAX_ENTITY_NUM = 512          # lib/features.py:37
MAX_SELECTED_UNITS_NUM = 64   # lib/features.py:36
EFFECT_LEN = 100              # lib/features.py:38
NUM_ACTIONS = 327

# (name, kind, width)   kind: 'o' one-hot(vocab=width), 'b' 11-bit binary MSB first, 'u' float unsqueeze
# actor_critic_default_config.yaml:264-364 (order is the concat order of entity_encoder.py:59-78)
ENTITY_FIELDS = [
    ('unit_type', 'o', 260), ('alliance', 'o', 5), ('cargo_space_taken', 'o', 9),
    ('build_progress', 'u', 1), ('health_ratio', 'u', 1), ('shield_ratio', 'u', 1), ('energy_ratio', 'u', 1),
    ('display_type', 'o', 5), ('x', 'b', 11), ('y', 'b', 11), ('cloak', 'o', 5), ('is_blip', 'o', 2),
    ('is_powered', 'o', 2), ('mineral_contents', 'u', 1), ('vespene_contents', 'u', 1),
    ('cargo_space_max', 'o', 9), ('assigned_harvesters', 'o', 24), ('weapon_cooldown', 'o', 32),
    ('order_length', 'o', 9), ('order_id_0', 'o', 327), ('order_id_1', 'o', 49), ('is_hallucination', 'o', 2),
    ('buff_id_0', 'o', 50), ('buff_id_1', 'o', 50), ('addon_unit_type', 'o', 9), ('is_active', 'o', 2),
    ('order_progress_0', 'u', 1), ('order_progress_1', 'u', 1), ('order_id_2', 'o', 49), ('order_id_3', 'o', 49),
    ('is_in_cargo', 'o', 2), ('attack_upgrade_level', 'o', 4), ('armor_upgrade_level', 'o', 4),
    ('shield_upgrade_level', 'o', 4), ('last_selected_units', 'o', 2), ('last_targeted_unit', 'o', 2),
]
ENTITY_INPUT_DIM = sum(w for _, _, w in ENTITY_FIELDS)
assert ENTITY_INPUT_DIM == 997

# (name, kind, in, out, scalar_context, baseline_feature)  scalar_encoder.py:99-132, yaml:146-219.
# 'time' sits 5th in the yaml but is skipped in the loop and appended last (scalar_encoder.py:106,127-128).
SCALAR_FIELDS = [
    ('agent_statistics', 'fc', 10, 64, False, True),
    ('home_race', 'emb', 5, 32, True, False),
    ('away_race', 'emb', 5, 32, True, False),
    ('upgrades', 'fc', 90, 128, False, True),
    ('unit_counts_bow', 'fc', 260, 128, False, True),
    ('last_delay', 'emb', 128, 64, False, False),
    ('last_queued', 'emb', 2, 32, False, False),
    ('last_action_type', 'emb', 327, 128, False, False),
    ('cumulative_stat', 'fc', 167, 128, True, True),
    ('beginning_order', 'bo', 214, 64, True, True),
    ('unit_type_bool', 'fc', 260, 64, True, False),
    ('enemy_unit_type_bool', 'fc', 260, 64, True, False),
    ('unit_order_type', 'fc', 269, 64, True, False),
]
TIME_DIM = 32
BO_LEN = 20
BO_ACTIONS = 174

# spatial planes, yaml:220-262 / spatial_encoder.py:51-71
SPATIAL_ONEHOT = [('visibility_map', 4), ('creep', 2), ('player_relative', 5), ('alerts', 2), ('pathable', 2),
                  ('buildable', 2)]
SPATIAL_EFFECTS = ['effect_PsiStorm', 'effect_NukeDot', 'effect_LiberatorDefenderZone', 'effect_BlindingCloud',
                   'effect_CorrosiveBile', 'effect_LurkerSpines']
HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']

# --------------------------------------------------------------------------------------------
# matmul operand rounding emulation (precision study only)
# --------------------------------------------------------------------------------------------
_EMU = 'fp32'
_EMU_LAYERS = None      # None: every product; else a tuple of layer-name prefixes the emulation is limited to


def set_matmul_emulation(mode: str, layers=None):
    """mode: operand rounding of the emulated products - 'fp32' (none), 'tf32', 'bf16' (both operands: the 1-term tensor-core
    product), 'bf16_w' / 'bf16_a' (only the weight / only the activation: what a 2-term product a.b_hi / a_hi.b keeps),
    'bf16x3' (the 3-term split, fp32-class).  layers: limit it to the layers whose parameter name starts with one of these
    prefixes ('@attention' selects the un-named attention products)."""
    global _EMU, _EMU_LAYERS
    assert mode in ('fp32', 'tf32', 'bf16', 'bf16x3', 'bf16_w', 'bf16_a')
    _EMU, _EMU_LAYERS = mode, (tuple(layers) if layers is not None else None)


def _rnd(x: Tensor, name: str = '', operand: str = 'a') -> Tensor:
    mode = _EMU
    if _EMU_LAYERS is not None and not name.startswith(_EMU_LAYERS):
        mode = 'fp32'
    if mode == 'bf16_w':
        mode = 'bf16' if operand == 'b' else 'fp32'
    elif mode == 'bf16_a':
        mode = 'bf16' if operand == 'a' else 'fp32'
    if mode == 'fp32' or mode == 'bf16x3':
        return x  # bf16x3 (hi*hi + hi*lo + lo*hi) keeps ~16 mantissa bits: treated as fp32 here
    if mode == 'bf16':
        return x.to(torch.bfloat16).to(torch.float32)
    # tf32: round-to-nearest-even onto a 10-bit mantissa
    i = x.contiguous().view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def _fc(P: Params, name: str, x: Tensor, relu: bool = False) -> Tensor:
    """fc_block(in,out[,ReLU]) — ctools/torch_utils/network/nn_module.py:231-270 (Linear at index 0)."""
    y = F.linear(_rnd(x, name, 'a'), _rnd(P[name + '.0.weight'], name, 'b'), P[name + '.0.bias'])
    return torch.relu(y) if relu else y


def _conv(P: Params, name: str, x: Tensor, pad: int, relu: bool = False) -> Tensor:
    """conv2d_block(...) — nn_module.py:119-174 (Conv2d at index 0, norm 'none')."""
    y = F.conv2d(_rnd(x, name, 'a'), _rnd(P[name + '.0.weight'], name, 'b'), P[name + '.0.bias'], padding=pad)
    return torch.relu(y) if relu else y


def _ln(P: Params, name: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), P[name + '.weight'], P[name + '.bias'], 1e-5)


def _mm(a: Tensor, b: Tensor, name: str = '@attention') -> Tensor:
    return torch.matmul(_rnd(a, name, 'a'), _rnd(b, name, 'b'))


def sequence_mask(lengths: Tensor, max_len: int) -> Tensor:
    """ctools/torch_utils/network/rnn.py:20-36."""
    return torch.arange(max_len, device=lengths.device).unsqueeze(0) < lengths.reshape(-1, 1)


def binary_rows(values: Tensor, bits: int) -> Tensor:
    """Frozen table of obs_encoder/entity_encoder.py:12-17: row n = bits of n, MSB first."""
    shifts = torch.arange(bits - 1, -1, -1, device=values.device)
    return ((values.long().unsqueeze(-1) >> shifts) & 1).float()


# --------------------------------------------------------------------------------------------
# Transformer (module_utils.py:71-199)
# --------------------------------------------------------------------------------------------
def _attention(P: Params, pre: str, x: Tensor, key_mask: Optional[Tensor], heads: int, head_dim: int) -> Tensor:
    """Attention.forward, module_utils.py:88-111.  key_mask [B,N] bool (True = valid key)."""
    B, N, _ = x.shape
    qkv = _fc(P, pre + '.attention_pre', x)
    q, k, v = torch.chunk(qkv, 3, dim=2)

    def split(t):
        return t.view(B, N, heads, head_dim).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    score = _mm(q, k.transpose(2, 3)) / math.sqrt(head_dim)
    if key_mask is not None:
        score = score.masked_fill(~key_mask.view(B, 1, 1, N), -1e9)
    p = torch.softmax(score, dim=-1)
    a = _mm(p, v).permute(0, 2, 1, 3).reshape(B, N, heads * head_dim)
    return _fc(P, pre + '.project', a)


def transformer(P: Params, pre: str, x: Tensor, key_mask: Optional[Tensor], heads: int, head_dim: int,
                layers: int, ln_type: str) -> Tensor:
    """Transformer.forward / TransformerLayer.forward, module_utils.py:130-151,191-199 (mlp_num=2, both ReLU)."""
    x = _fc(P, pre + '.embedding', x, relu=True)
    for i in range(layers):
        lp = '%s.layers.%d' % (pre, i)
        if ln_type == 'post':
            a = _attention(P, lp + '.attention', x, key_mask, heads, head_dim)
            x = _ln(P, lp + '.layernorm1', x + a)
            m = _fc(P, lp + '.mlp.1', _fc(P, lp + '.mlp.0', x, relu=True), relu=True)
            x = _ln(P, lp + '.layernorm2', x + m)
        else:
            a = _attention(P, lp + '.attention', _ln(P, lp + '.layernorm1', x), key_mask, heads, head_dim)
            x = x + a
            m = _fc(P, lp + '.mlp.1', _fc(P, lp + '.mlp.0', _ln(P, lp + '.layernorm2', x), relu=True), relu=True)
            x = x + m
    return x


# --------------------------------------------------------------------------------------------
# Encoders
# --------------------------------------------------------------------------------------------
def scalar_encoder(P: Params, s: Dict[str, Tensor], spatial_x: int) -> Tuple[Tensor, Tensor, Tensor]:
    """ScalarEncoder.forward, obs_encoder/scalar_encoder.py:99-132 (+ BeginningBuildOrderEncoder :42-53)."""
    pre = 'encoder.scalar_encoder.encode_modules.'
    outs, ctx, base = [], [], []
    for name, kind, din, dout, is_ctx, is_base in SCALAR_FIELDS:
        if kind == 'emb':
            idx = s[name].long().clamp(max=din - 1)
            e = torch.relu(P[pre + name + '.weight'][idx])
        elif kind == 'fc':
            e = _fc(P, pre + name, s[name].float(), relu=True)
        else:
            bo = s['beginning_order'].long()
            loc = s['bo_location'].long()
            B = bo.shape[0]
            tok = torch.cat([
                F.one_hot(bo, BO_ACTIONS).float(),
                torch.eye(BO_LEN).unsqueeze(0).expand(B, -1, -1),
                binary_rows(loc % spatial_x, 10), binary_rows(loc // spatial_x, 10)], dim=2)
            t = transformer(P, pre + 'beginning_order.transformer', tok, None, heads=2, head_dim=8, layers=3,
                            ln_type='pre')
            e = _fc(P, pre + 'beginning_order.embedd_fc', t.mean(dim=1), relu=True)
        outs.append(e)
        if is_ctx:
            ctx.append(e)
        if is_base:
            base.append(e)
    # time_encoder, scalar_encoder.py:91-97 with compute_denominator :11-16
    pa = P['encoder.scalar_encoder.position_array']
    t = s['time'].float().unsqueeze(1)
    te = torch.zeros(t.shape[0], TIME_DIM)
    te[:, 0::2] = torch.sin(t * pa[0::2])
    te[:, 1::2] = torch.cos(t * pa[1::2])
    outs.append(te)
    return torch.cat(outs, 1), torch.cat(ctx, 1), torch.cat(base, 1)


def entity_features(e: Dict[str, Tensor]) -> Tensor:
    """The 997-wide concat of entity_encoder.py:59-78 (one-hot ids >= vocab clamp, negatives raise)."""
    cols = []
    for name, kind, w in ENTITY_FIELDS:
        v = e[name]
        if kind == 'o':
            if (v < 0).any():
                raise RuntimeError('negative categorical id in entity field %s' % name)
            cols.append(F.one_hot(v.long().clamp(max=w - 1), w).float())
        elif kind == 'b':
            cols.append(binary_rows(v, w))
        else:
            cols.append(v.float().unsqueeze(-1))
    return torch.cat(cols, dim=-1)


def entity_encoder(P: Params, e: Dict[str, Tensor], entity_num: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """EntityEncoder.forward, obs_encoder/entity_encoder.py:59-96, entity_reduce_type 'selected_units_num'.

    Quirk kept: `self.act(x)` is the shared in-place ReLU (:39,:81) so the pooled mean is over relu(x).
    """
    pre = 'encoder.entity_encoder.'
    x = entity_features(e)
    mask = sequence_mask(entity_num, x.shape[1])
    x = transformer(P, pre + 'transformer', x, mask, heads=2, head_dim=128, layers=3, ln_type='post')
    x = torch.relu(x)
    entity_embeddings = _fc(P, pre + 'entity_fc', x, relu=True)
    pooled = (x * mask.unsqueeze(2)).sum(dim=1) / entity_num.unsqueeze(-1)
    embedded_entity = _fc(P, pre + 'embed_fc', pooled, relu=True)
    return entity_embeddings, embedded_entity, mask


def scatter_connection(project: Tensor, x: Tensor, y: Tensor, H: int, W: int) -> Tensor:
    """scatter_connection(..., 'add'), module_utils.py:11-34: map[n,c,clamp(y),clamp(x)] += project[n,e,c]."""
    N, E, C = project.shape
    idx = y.long().clamp(0, H - 1) * W + x.long().clamp(0, W - 1)           # [N,E]
    out = torch.zeros(N, H * W, C)
    out.scatter_add_(1, idx.unsqueeze(-1).expand(-1, -1, C), project)
    return out.view(N, H, W, C).permute(0, 3, 1, 2)


def spatial_planes(sp: Dict[str, Tensor], scatter_map: Tensor) -> Tensor:
    """The 56-channel concat of spatial_encoder.py:51-71.

    Quirk kept: effect lists are zero padded, so flat pixel 0 of every effect plane is always 1 (:62-69).
    """
    N, H, W = sp['height_map'].shape
    planes = [sp['height_map'].float().unsqueeze(1) / 256]
    for name, n in SPATIAL_ONEHOT:
        planes.append(F.one_hot(sp[name].long(), n).float().permute(0, 3, 1, 2))
    for name in SPATIAL_EFFECTS:
        p = torch.zeros(N, H * W)
        p.scatter_(1, sp[name].long(), 1.0)
        planes.append(p.view(N, 1, H, W))
    planes.append(scatter_map)
    return torch.cat(planes, dim=1)


def spatial_encoder(P: Params, sp: Dict[str, Tensor], scatter_map: Tensor) -> Tuple[Tensor, List[Tensor]]:
    """SpatialEncoder.forward, obs_encoder/spatial_encoder.py:51-90 (maxpool downsample, 4 ResBlocks, fc head)."""
    pre = 'encoder.spatial_encoder.'
    x = _conv(P, pre + 'project', spatial_planes(sp, scatter_map), 0, relu=True)
    skips = []
    for i in range(3):
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
        x = _conv(P, pre + 'downsample.%d' % i, x, 1, relu=True)
    for i in range(4):
        skips.append(x)
        r = _conv(P, pre + 'res.%d.conv1' % i, x, 1, relu=True)      # res_block.py:56-65
        r = _conv(P, pre + 'res.%d.conv2' % i, r, 1)
        x = torch.relu(r + x)
    x = _fc(P, pre + 'fc', x.reshape(x.shape[0], -1), relu=True)
    return x, skips


def encoder(P: Params, spatial_info, entity_info, scalar_info, entity_num):
    """Encoder.forward, model/encoder.py:28-45."""
    N, H, W = spatial_info['height_map'].shape
    embedded_scalar, scalar_context, baseline_feature = scalar_encoder(P, scalar_info, W)
    entity_embeddings, embedded_entity, mask = entity_encoder(P, entity_info, entity_num)
    project = _fc(P, 'encoder.scatter_project', entity_embeddings, relu=True) * mask.unsqueeze(2)
    scatter_map = scatter_connection(project, entity_info['x'], entity_info['y'], H, W)
    embedded_spatial, map_skip = spatial_encoder(P, spatial_info, scatter_map)
    lstm_input = torch.cat([embedded_scalar, embedded_entity, embedded_spatial], dim=-1)
    return lstm_input, scalar_context, baseline_feature, entity_embeddings, map_skip


VALUE_FC_FIELDS = ['enemy_unit_counts_bow', 'enemy_unit_type_bool', 'enemy_agent_statistics', 'enemy_upgrades', 'cumulative_stat']


def _bo_tokens(bo: Tensor, loc: Tensor, spatial_x: int) -> Tensor:
    """BeginningBuildOrderEncoder.forward up to the transformer input, scalar_encoder.py:34-48."""
    B = bo.shape[0]
    return torch.cat([F.one_hot(bo.long(), BO_ACTIONS).float(), torch.eye(BO_LEN).unsqueeze(0).expand(B, -1, -1),
                      binary_rows(loc.long() % spatial_x, 10), binary_rows(loc.long() // spatial_x, 10)], dim=2)


def value_encoder(P: Params, vf: Dict[str, Tensor], spatial_x: int) -> Tensor:
    """ValueEncoder.forward, obs_encoder/value_encoder.py:47-74 (module order: actor_critic_default_config.yaml:29-64)."""
    pre = 'value_encoder.'
    em = pre + 'encode_modules.'
    fc = [_fc(P, em + k, vf[k].float(), relu=True) for k in VALUE_FC_FIELDS]                       # :50-53
    unit = torch.cat([P[em + 'unit_alliance.weight'][vf['unit_alliance'].long()],
                      P[em + 'unit_type.weight'][vf['unit_type'].long()]], dim=-1)                  # :54-56 (nn.Embedding)
    t = transformer(P, em + 'beginning_order.transformer', _bo_tokens(vf['beginning_order'], vf['bo_location'], spatial_x),
                    None, heads=2, head_dim=8, layers=3, ln_type='pre')
    bo = _fc(P, em + 'beginning_order.embedd_fc', t.mean(dim=1), relu=True)                         # :58
    project = _fc(P, pre + 'scatter_project', unit, relu=True)                                      # :61
    project = project * sequence_mask(vf['total_unit_count'], project.shape[1]).unsqueeze(2)        # :62-63
    N, _c, H, W = vf['own_units_spatial'].shape
    scatter_map = scatter_connection(project, vf['unit_x'], vf['unit_y'], H, W)                     # :64-66
    x = torch.cat([scatter_map, vf['own_units_spatial'].float(), vf['enemy_units_spatial'].float()], dim=1)
    x = _conv(P, pre + 'project', x, 0, relu=True)                                                  # :68
    for i in range(3):                                                                              # :69 nn.Sequential(pool, conv) x 3
        x = _conv(P, pre + 'downsample.%d' % (2 * i + 1), F.max_pool2d(x, 2, 2), 1, relu=True)
    for i in range(4):                                                                              # :70-71 ResBlock, res_block.py:56-65
        r = _conv(P, pre + 'res.%d.conv1' % i, x, 1, relu=True)
        x = torch.relu(_conv(P, pre + 'res.%d.conv2' % i, r, 1) + x)
    x = _fc(P, pre + 'spatial_fc', x.reshape(N, -1), relu=True)                                     # :72
    return torch.cat(fc + [x, bo], dim=-1)                                                          # :73


# --------------------------------------------------------------------------------------------
# LayerNorm LSTM (model/lstm.py:120-167,215-234)
# --------------------------------------------------------------------------------------------
def lnlstm_cell(P: Params, pre: str, x: Tensor, h: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
    """LayerNormLSTMCell.forward, lstm.py:138-153; gate order in/forget/cell/out; carried c is layer-normed."""
    ig = _ln(P, pre + '.layernorm_i', _mm(x, P[pre + '.weight_ih'].t(), pre + '.weight_ih'))
    hg = _ln(P, pre + '.layernorm_h', _mm(h, P[pre + '.weight_hh'].t(), pre + '.weight_hh'))
    i, f, g, o = (ig + hg).chunk(4, 1)
    c2 = _ln(P, pre + '.layernorm_c', torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g))
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def lnlstm(P: Params, pre: str, x: Tensor, state: List[Tuple[Tensor, Tensor]], layers: int):
    """StackedLSTM over [L,B,D], lstm.py:161-167,223-234 (layer-major loop order)."""
    out_state = []
    for l in range(layers):
        h, c = state[l]
        ys = []
        for t in range(x.shape[0]):
            h, c = lnlstm_cell(P, '%s.layers.%d.cell' % (pre, l), x[t], h, c)
            ys.append(h)
        x = torch.stack(ys)
        out_state.append((h, c))
    return x, out_state


# --------------------------------------------------------------------------------------------
# Sampling
# --------------------------------------------------------------------------------------------
def sample_categorical(logits: Tensor) -> Tensor:
    """`torch.multinomial(F.softmax(x, 1), 1)[:, 0]` as used at action_type_head.py:57-58 etc.

    Calls the same ATen op so the default-generator stream is consumed exactly like the reference.
    (n=1 fast path == argmax(p / q), q ~ Exp(1) drawn with p's shape; checked in tests.)
    """
    return torch.multinomial(torch.softmax(logits, dim=-1), 1)[:, 0]


# --------------------------------------------------------------------------------------------
# Heads
# --------------------------------------------------------------------------------------------
def _glu(P: Params, name: str, x: Tensor, context: Tensor) -> Tensor:
    """GLU.forward, module_utils.py:519-524."""
    return _fc(P, name + '.layer2', torch.sigmoid(_fc(P, name + '.layer1', context)) * x)


def action_type_head(P: Params, lstm_out: Tensor, scalar_context: Tensor, temperature: float,
                     action_type: Optional[Tensor] = None):
    """ActionTypeHead.forward, head/action_type_head.py:48-67 (use_mask False outside 'play')."""
    pre = 'policy.action_type_head.'
    x = _fc(P, pre + 'project', lstm_out, relu=True)
    for i in range(2):                                                 # ResFCBlock(LN), res_block.py:68-108
        r = torch.relu(_ln(P, pre + 'res.%d.fc1.1' % i, _fc(P, pre + 'res.%d.fc1' % i, x)))
        r = _ln(P, pre + 'res.%d.fc2.1' % i, _fc(P, pre + 'res.%d.fc2' % i, r))
        x = torch.relu(r + x)
    logits = _glu(P, pre + 'action_fc', x, scalar_context) / temperature
    if action_type is None:
        action_type = sample_categorical(logits)
    one_hot = F.one_hot(action_type.long(), NUM_ACTIONS).float()
    e1 = _fc(P, pre + 'action_map_fc2', _fc(P, pre + 'action_map_fc1', one_hot, relu=True))
    e1 = _glu(P, pre + 'glu1', e1, scalar_context)
    e2 = _glu(P, pre + 'glu2', lstm_out, scalar_context)
    return logits, action_type, e1 + e2


def _arg_head(P: Params, pre: str, emb: Tensor, n: int, temperature: Optional[float], action: Optional[Tensor]):
    """DelayHead / QueuedHead, head/action_arg_head.py:41-53,73-86 (delay logits are NOT divided by T)."""
    x = _fc(P, pre + 'fc3', _fc(P, pre + 'fc2', _fc(P, pre + 'fc1', emb, relu=True), relu=True))
    if temperature is not None:
        x = x / temperature
    if action is None:
        action = sample_categorical(x)
    e = _fc(P, pre + 'embed_fc2', _fc(P, pre + 'embed_fc1', F.one_hot(action.long(), n).float(), relu=True))
    return x, action, emb + e


def _su_keys(P: Params, entity_embeddings: Tensor, entity_num: Tensor):
    """SelectedUnitsHead._get_key_mask, action_arg_head.py:118-143: learned end token at slot entity_num."""
    pre = 'policy.selected_units_head.'
    N, E, _ = entity_embeddings.shape
    key = torch.cat([_fc(P, pre + 'key_fc', entity_embeddings), torch.zeros(N, 1, 32)], dim=1)
    rows = torch.arange(N)
    key = key.clone()
    key[rows, entity_num] = P[pre + 'end_embedding'][0]
    mask = sequence_mask(entity_num + 1, E + 1)
    return key, mask


def _su_embed(P: Params, key: Tensor, one_hot: Tensor, normalise: Tensor) -> Tensor:
    """Masked mean of selected keys -> embed_fc2(embed_fc1(.)), action_arg_head.py:196-199 / :290-293."""
    pre = 'policy.selected_units_head.'
    s = (key * one_hot.unsqueeze(2)).sum(dim=1)
    cnt = one_hot.sum(dim=1, keepdim=True)
    s = torch.where(normalise.unsqueeze(1), s / cnt, s)
    return _fc(P, pre + 'embed_fc2', _fc(P, pre + 'embed_fc1', s, relu=True))


def selected_units_head_train(P: Params, emb0: Tensor, entity_embeddings: Tensor, entity_num: Tensor,
                              selected_units_num: Tensor, selected_units: Tensor, temperature: float):
    """Teacher-forced SelectedUnitsHead._query, action_arg_head.py:168-216.

    Quirks kept: loop length = max(selected_units_num) over the whole flattened batch (min 1); `end_flag`
    persists; rows with selected_units_num == 0 are not normalised (:197-198) and rows whose count is 0
    with num != 0 divide 0/0 (only num == 1, never produced by the sampler).  Train logits are NOT divided
    by the temperature (only the sampling path calls _get_pred_with_logit).
    """
    pre = 'policy.selected_units_head.'
    N = emb0.shape[0]
    rows = torch.arange(N)
    key, mask = _su_keys(P, entity_embeddings, entity_num)
    base_mask = mask.clone()
    base_mask[rows, entity_num] = False                                 # :163 and :180
    S = max(int(selected_units_num.max()), 1)
    end_flag = torch.zeros(N, dtype=torch.bool)
    one_hot = torch.zeros(N, key.shape[1])
    h = torch.zeros(N, 32)
    c = torch.zeros(N, 32)
    ae = emb0
    step_mask = base_mask
    logits = []
    for i in range(S):
        if i > 0:
            step_mask = step_mask.clone()
            if i == 1:
                step_mask[rows, entity_num] = True
            step_mask[rows, selected_units[:, i - 1]] = False
        q = _fc(P, pre + 'query_fc2', _fc(P, pre + 'query_fc1', ae, relu=True))
        h, c = lnlstm_cell(P, pre + 'lstm.layers.0.cell', q, h, c)
        logits.append(((h.unsqueeze(1) * key).sum(dim=2)).masked_fill(~step_mask, -1e9))
        end_flag = end_flag | (selected_units[:, i] == entity_num)
        one_hot = one_hot.clone()
        sel = ~end_flag
        one_hot[rows[sel], selected_units[:, i][sel]] = 1
        ae = emb0 + _su_embed(P, key, one_hot, selected_units_num != 0)
    # the reference returns `results` = None on the teacher-forced path (action_arg_head.py:166,314)
    return torch.stack(logits, dim=1), None, ae, selected_units_num


def selected_units_head_sample(P: Params, emb0: Tensor, entity_embeddings: Tensor, entity_num: Tensor,
                               su_mask: Tensor, temperature: float):
    """Sampling SelectedUnitsHead._query, action_arg_head.py:262-314 (extra_units disabled -> zeros)."""
    pre = 'policy.selected_units_head.'
    N = emb0.shape[0]
    rows = torch.arange(N)
    key, mask = _su_keys(P, entity_embeddings, entity_num)
    step_mask = mask.clone()
    step_mask[rows, entity_num] = False
    num = torch.full((N,), MAX_SELECTED_UNITS_NUM, dtype=torch.long)
    end_flag = ~su_mask.clone()
    num[~su_mask] = 0
    one_hot = torch.zeros(N, key.shape[1])
    h = torch.zeros(N, 32)
    c = torch.zeros(N, 32)
    ae = emb0
    results, logits = [], []
    result = None
    for i in range(MAX_SELECTED_UNITS_NUM):
        if i > 0:
            if i == 1:
                step_mask[rows, entity_num] = True
            step_mask[rows, result] = False
        q = _fc(P, pre + 'query_fc2', _fc(P, pre + 'query_fc1', ae, relu=True))
        h, c = lnlstm_cell(P, pre + 'lstm.layers.0.cell', q, h, c)
        step_logits = ((h.unsqueeze(1) * key).sum(dim=2)).masked_fill(~step_mask, -1e9) / temperature
        result = sample_categorical(step_logits)
        num[(result == entity_num) & ~end_flag] = i + 1
        end_flag = end_flag | (result == entity_num)
        results.append(result)
        logits.append(step_logits)
        sel = ~end_flag
        one_hot[rows[sel], result[sel]] = 1
        cnt = one_hot.sum(dim=1)
        ae = emb0 + _su_embed(P, key, one_hot, cnt != 0)
        if bool(end_flag.all()):
            break
    extra_units = torch.zeros(N, MAX_ENTITY_NUM + 1)
    return torch.stack(logits, dim=1), torch.stack(results, dim=1), ae, num, extra_units


def target_unit_head(P: Params, emb: Tensor, entity_embeddings: Tensor, entity_num: Tensor, temperature: float,
                     target_unit: Optional[Tensor] = None):
    """TargetUnitHead.forward, action_arg_head.py:343-363."""
    pre = 'policy.target_unit_head.'
    key = _fc(P, pre + 'key_fc', entity_embeddings)
    q = _fc(P, pre + 'query_fc2', _fc(P, pre + 'query_fc1', emb, relu=True))
    logits = (q.unsqueeze(1) * key).sum(dim=2)
    logits = logits.masked_fill(~sequence_mask(entity_num, entity_embeddings.shape[1]), -1e9) / temperature
    if target_unit is None:
        target_unit = sample_categorical(logits)
    return logits, target_unit


def location_head(P: Params, emb: Tensor, map_skip: List[Tensor], temperature: float,
                  location: Optional[Tensor] = None):
    """LocationHead.forward, action_arg_head.py:417-450 (gate=True, film/unet False, bilinear upsample)."""
    pre = 'policy.location_head.'
    N = emb.shape[0]
    h8, w8 = map_skip[-1].shape[2:]
    x = _fc(P, pre + 'project_embed', emb, relu=True).reshape(N, 4, h8, w8)
    x = torch.relu(torch.cat([x, map_skip[-1]], dim=1))
    x = _conv(P, pre + 'conv1', x, 0, relu=True)
    for i in range(4):
        x = x + map_skip[len(map_skip) - i - 1]
        rp = pre + 'res.%d.' % i                                       # GatedResBlock, module_utils.py:224-231
        r = _conv(P, rp + 'conv2', _conv(P, rp + 'conv1', x, 1, relu=True), 1)
        g = x
        for j in range(4):
            g = _conv(P, rp + 'GateWeightG.%d' % j, g, 0, relu=(j < 3))
        r = torch.tanh(r * torch.sigmoid(g)) * P[rp + 'UpdateSP']
        x = torch.relu(r + x)
    for i in range(3):
        x = F.interpolate(x, scale_factor=2., mode='bilinear')
        x = _conv(P, pre + 'upsample.%d' % i, x, 1, relu=(i < 2))
    logits = x.reshape(N, -1) / temperature
    if location is None:
        location = sample_categorical(logits)
    return logits, location


def value_baseline(P: Params, name: str, x: Tensor, atan: bool) -> Tensor:
    """ValueBaseline.forward, model/value.py:31-39 with ResFCBlock2, res_block.py:110-141."""
    pre = 'value_networks.%s.' % name
    x = _fc(P, pre + 'project', x, relu=True)
    for i in range(16):
        r = _fc(P, pre + 'res.%d.fc2' % i, _fc(P, pre + 'res.%d.fc1' % i, x, relu=True))
        x = _ln(P, pre + 'res.%d.norm' % i, r + x)
    v = _fc(P, pre + 'value_fc', x).squeeze(1)
    if atan:
        v = (2.0 / math.pi) * torch.atan((math.pi / 2.0) * v)
    return v


# --------------------------------------------------------------------------------------------
# Policy and Model entry points
# --------------------------------------------------------------------------------------------
def policy_sample(P: Params, lstm_out, entity_embeddings, map_skip, scalar_context, entity_num,
                  su_action_mask: Tensor, temperature: float = 1.0):
    """Policy.forward, model/policy.py:22-48.  su_action_mask = SELECTED_UNITS_MASK table [327] bool."""
    logit, action = {}, {}
    logit['action_type'], action['action_type'], emb = action_type_head(P, lstm_out, scalar_context, temperature)
    logit['delay'], action['delay'], emb = _arg_head(P, 'policy.delay_head.', emb, 128, None, None)
    logit['queued'], action['queued'], emb = _arg_head(P, 'policy.queued_head.', emb, 2, temperature, None)
    su_mask = su_action_mask[action['action_type']]
    logit['selected_units'], action['selected_units'], emb, su_num, extra = selected_units_head_sample(
        P, emb, entity_embeddings, entity_num, su_mask, temperature)
    logit['target_unit'], action['target_unit'] = target_unit_head(P, emb, entity_embeddings, entity_num,
                                                                   temperature)
    logit['target_location'], action['target_location'] = location_head(P, emb, map_skip, temperature)
    return action, su_num, logit, extra


def policy_train(P: Params, lstm_out, entity_embeddings, map_skip, scalar_context, entity_num, action_info,
                 selected_units_num, temperature: float = 1.0):
    """Policy.train_forward, model/policy.py:50-73."""
    logit, action = {}, {}
    logit['action_type'], action['action_type'], emb = action_type_head(
        P, lstm_out, scalar_context, temperature, action_info['action_type'])
    logit['delay'], action['delay'], emb = _arg_head(P, 'policy.delay_head.', emb, 128, None, action_info['delay'])
    logit['queued'], action['queued'], emb = _arg_head(P, 'policy.queued_head.', emb, 2, temperature,
                                                       action_info['queued'])
    logit['selected_units'], action['selected_units'], emb, su_num = selected_units_head_train(
        P, emb, entity_embeddings, entity_num, selected_units_num, action_info['selected_units'], temperature)
    logit['target_unit'], action['target_unit'] = target_unit_head(
        P, emb, entity_embeddings, entity_num, temperature, action_info['target_unit'])
    logit['target_location'], action['target_location'] = location_head(
        P, emb, map_skip, temperature, action_info['target_location'])
    return action, su_num, logit


def compute_logp_action(P: Params, spatial_info, entity_info, scalar_info, entity_num, hidden_state,
                        su_action_mask: Tensor, temperature: float = 1.0, **_):
    """Model.compute_logp_action, model/model.py:56-74."""
    lstm_input, scalar_context, _b, entity_embeddings, map_skip = encoder(
        P, spatial_info, entity_info, scalar_info, entity_num)
    lstm_out, out_state = lnlstm(P, 'core_lstm', lstm_input.unsqueeze(0), hidden_state, 3)
    action, su_num, logit, extra = policy_sample(P, lstm_out.squeeze(0), entity_embeddings, map_skip,
                                                 scalar_context, entity_num, su_action_mask, temperature)
    logp = {}
    for k, a in action.items():
        lp = torch.log_softmax(logit[k], dim=-1)
        logp[k] = lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)
    return {'action_info': action, 'action_logp': logp, 'selected_units_num': su_num, 'entity_num': entity_num,
            'hidden_state': out_state, 'logit': logit, 'extra_units': extra}


def compute_teacher_logit(P: Params, spatial_info, entity_info, scalar_info, entity_num, hidden_state,
                          selected_units_num, action_info, temperature: float = 1.0, **_):
    """Model.compute_teacher_logit, model/model.py:76-93."""
    lstm_input, scalar_context, _b, entity_embeddings, map_skip = encoder(
        P, spatial_info, entity_info, scalar_info, entity_num)
    lstm_out, out_state = lnlstm(P, 'core_lstm', lstm_input.unsqueeze(0), hidden_state, 3)
    _a, su_num, logit = policy_train(P, lstm_out.squeeze(0), entity_embeddings, map_skip, scalar_context,
                                     entity_num, action_info, selected_units_num, temperature)
    return {'logit': logit, 'hidden_state': out_state, 'entity_num': entity_num, 'selected_units_num': su_num}


BASELINE_ATAN = {'winloss': True, 'build_order': False, 'built_unit': False, 'effect': False, 'upgrade': False,
                 'battle': False}


def enabled_baselines(P: Params) -> List[str]:
    return [k for k in BASELINE_ATAN if ('value_networks.%s.project.0.weight' % k) in P]


def rl_learner_forward(P: Params, spatial_info, entity_info, scalar_info, entity_num, hidden_state, action_info,
                       selected_units_num, behaviour_logp, teacher_logit, mask, reward, step, batch_size,
                       unroll_len, temperature: float = 1.0, value_feature=None, **_):
    """Model.rl_learner_forward, model/model.py:95-168 (only_update_baseline False; use_value_feature when the weights hold
    a value encoder: critic input = [lstm output | value feature | scalar encoder baseline feature], :141-144)."""
    B, T = batch_size, unroll_len
    flat_action = {k: v.flatten(0, 1) for k, v in action_info.items()}
    flat_su_num = selected_units_num.flatten(0, 1)
    lstm_input, scalar_context, baseline_feature, entity_embeddings, map_skip = encoder(
        P, spatial_info, entity_info, scalar_info, entity_num)
    state0 = [(h.view(-1, B, h.shape[-1])[0], c.view(-1, B, c.shape[-1])[0]) for h, c in hidden_state]
    lstm_out, _ = lnlstm(P, 'core_lstm', lstm_input.view(-1, B, lstm_input.shape[-1]), state0, 3)
    lstm_out = lstm_out.reshape(-1, lstm_out.shape[-1])
    _a, _n, logits = policy_train(P, lstm_out[:-B], entity_embeddings[:-B], [m[:-B] for m in map_skip],
                                  scalar_context[:-B], entity_num[:-B], flat_action, flat_su_num, temperature)
    critic_input = lstm_out
    if 'value_encoder.project.0.weight' in P:
        W = spatial_info['height_map'].shape[-1]
        critic_input = torch.cat([lstm_out, value_encoder(P, value_feature, W), baseline_feature], dim=1)
    values = {k: value_baseline(P, k, critic_input, BASELINE_ATAN[k]).view(T + 1, B) for k in enabled_baselines(P)}
    logits = {k: v.view(T, B, *v.shape[1:]) for k, v in logits.items()}
    su = logits['selected_units']
    logits['selected_units'] = F.pad(su, (0, 0, 0, MAX_SELECTED_UNITS_NUM - su.shape[2]), 'constant', -1e9)
    return {'unroll_len': T, 'batch_size': B, 'selected_units_num': selected_units_num, 'target_logit': logits,
            'value': values, 'action_log_prob': behaviour_logp, 'teacher_logit': teacher_logit, 'mask': mask,
            'action': action_info, 'reward': reward, 'step': step}


def sl_train(P: Params, spatial_info, entity_info, scalar_info, entity_num, selected_units_num, traj_lens,
             hidden_state, action_info, temperature: float = 1.0, **_):
    """Model.sl_train, model/model.py:170-189 (obs rows batch-major [B*T])."""
    B = len(traj_lens)
    lstm_input, scalar_context, _b, entity_embeddings, map_skip = encoder(
        P, spatial_info, entity_info, scalar_info, entity_num)
    x = lstm_input.view(-1, lstm_input.shape[0] // B, lstm_input.shape[-1]).permute(1, 0, 2)
    lstm_out, out_state = lnlstm(P, 'core_lstm', x, hidden_state, 3)
    lstm_out = lstm_out.permute(1, 0, 2).reshape(-1, lstm_out.shape[-1])
    action, su_num, logits = policy_train(P, lstm_out, entity_embeddings, map_skip, scalar_context, entity_num,
                                          action_info, selected_units_num, temperature)
    return logits, action, out_state


# --------------------------------------------------------------------------------------------
# RL loss (rl_training/rl_loss.py:33-185, rl_training/as_rl_utils.py)
# --------------------------------------------------------------------------------------------
def vtrace_advantages(rho: Tensor, reward: Tensor, value: Tensor) -> Tensor:
    """as_rl_utils.py:284-312 with gamma=1, lambda=1, c = rho (call site :15)."""
    T = reward.shape[0]
    delta = rho * (reward + value[1:] - value[:-1])
    vs = torch.empty_like(value)
    vs[-1] = value[-1]
    for t in reversed(range(T)):
        vs[t] = value[t] + delta[t] + rho[t] * (vs[t + 1] - value[t + 1])
    return rho * (reward + vs[1:] - value[:-1])


def lambda_returns(reward: Tensor, value: Tensor, gamma: float, lam) -> Tensor:
    """generalized_lambda_returns / multistep_forward_view, as_rl_utils.py:157-218."""
    T = reward.shape[0]
    lam = lam if isinstance(lam, Tensor) else torch.full_like(reward, lam)
    out = torch.empty_like(reward)
    out[-1] = reward[-1] + gamma * value[-1]
    for t in reversed(range(T - 1)):
        out[t] = reward[t] + gamma * lam[t] * out[t + 1] + gamma * (1 - lam[t]) * value[t + 1]
    return out


def upgo_returns(reward: Tensor, value: Tensor) -> Tensor:
    """as_rl_utils.py:265-281 (note the >=)."""
    k = ((reward + value[1:]) >= value[:-1]).float()
    lam = torch.cat([k[1:], torch.ones_like(k[-1:])], dim=0)
    return lambda_returns(reward, value, 1.0, lam)


DEFAULT_RL_LOSS_CFG = {  # bin/rl_user_config.yaml:58-117 merged over rl_training/default_reinforcement_loss.yaml
    'baseline_w': {'winloss': 10.0, 'build_order': 0.0, 'built_unit': 0.0, 'effect': 0.0, 'upgrade': 0.0,
                   'battle': 0.0},
    'pg_w': {'winloss': 1.0, 'build_order': 0.0, 'built_unit': 0.0, 'effect': 0.0, 'upgrade': 0.0, 'battle': 0.0},
    'upgo_w': 1.0, 'kl_w': 0.002, 'action_type_kl_w': 0.1, 'entropy_w': 0.0001,
    'head_w': {h: 1.0 for h in HEADS},      # pg / upgo / entropy / kl head weights are all 1 in the yaml
    'action_type_kl_steps': 5200,
    'gamma_baseline': {'winloss': 1.0, 'build_order': 1.0, 'built_unit': 1.0, 'effect': 1.0, 'upgrade': 1.0,
                       'battle': 0.997},
}


def rl_loss(out: dict, cfg: dict = None, only_update_value: bool = False, use_dapo: bool = False,
            dapo_w: float = 0.1, dapo_steps: int = 2400) -> Dict[str, Tensor]:
    """ReinforcementLoss.compute_loss, rl_loss.py:33-185.  Returns tensors (no .item()).  use_dapo: the 'MP' players' extra
    KL(successive || target) term (rl_loss.py:164-172, as_rl_utils.py:105-127) over out['successive_logit']."""
    cfg = cfg or DEFAULT_RL_LOSS_CFG
    logits, values = out['target_logit'], out['value']          # zeroing below is visible to the caller, as in the reference
    mu, teacher, mask, action, reward, step = (out['action_log_prob'], out['teacher_logit'], out['mask'],
                                               out['action'], out['reward'], out['step'])
    info = {}
    flag = (reward['winloss'][-1] == 0)
    for f in values:                                                     # rl_loss.py:47-49
        v = values[f].clone()
        v[-1] = v[-1] * flag
        values[f] = v
    su_mask = mask['selected_units_mask']
    logp_all, prob_all, lam, rho = {}, {}, {}, {}
    for h in HEADS:                                                      # rl_loss.py:63-90
        lp = torch.log_softmax(logits[h], dim=-1)
        logp_all[h], prob_all[h] = lp, lp.exp()
        la = lp.gather(-1, action[h].unsqueeze(-1)).squeeze(-1)
        with torch.no_grad():
            lr = la - mu[h]
            if h == 'selected_units':
                lr = (lr * su_mask).sum(-1)
            rho[h] = lr.exp().clamp(max=1)
        if h == 'selected_units':
            la = la.masked_fill(~su_mask, 0).sum(-1)
        lam[h] = la

    def am(h):
        return 1.0 if h in ('action_type', 'delay') else mask['actions_mask'][h]

    total_pg = 0.
    for f, v in values.items():                                          # as_rl_utils.py:1-28
        tot = 0.
        for h in HEADS:
            with torch.no_grad():
                adv = vtrace_advantages(rho[h], reward[f], v)
            l = -adv * lam[h] * am(h)
            if f in ('build_order', 'built_unit', 'effect'):
                l = l * mask[f + '_mask']
            l = l.mean()
            info['%s/%s' % (f, h)] = l.detach()
            tot = tot + l * cfg['head_w'][h]
        info[f + '/total'] = tot.detach()
        total_pg = total_pg + cfg['pg_w'][f] * tot
    total_upgo = 0.                                                      # as_rl_utils.py:31-49
    with torch.no_grad():
        ret = upgo_returns(reward['winloss'], values['winloss'])
    for h in HEADS:
        with torch.no_grad():
            adv = rho[h] * (ret - values['winloss'][:-1])
        l = (-adv * lam[h] * am(h)).mean()
        info['upgo/' + h] = l.detach()
        total_upgo = total_upgo + l * cfg['head_w'][h]
    info['upgo/total'] = total_upgo.detach()
    total_upgo = total_upgo * cfg['upgo_w']
    total_critic = 0.                                                    # as_rl_utils.py:221-243
    for f, v in values.items():
        with torch.no_grad():
            ret = lambda_returns(reward[f], v, cfg['gamma_baseline'][f], 0.8)
        l = 0.5 * (ret - v[:-1]) ** 2
        if f in ('build_order', 'built_unit', 'effect'):
            l = l * mask[f + '_mask']
        l = l.mean()
        total_critic = total_critic + cfg['baseline_w'][f] * l
        info[f + '/td'] = l.detach()
        info[f + '/reward'] = reward[f].float().mean()
        info[f + '/value'] = v.mean().detach()
    info['battle/reward'] = reward['battle'].float().mean()
    total_ent = 0.                                                       # as_rl_utils.py:52-72
    for h in HEADS:
        ent = -(prob_all[h] * logp_all[h]).sum(-1)
        if h == 'selected_units':
            ent = ent / (1e-9 + torch.log(mask['selected_units_logits_mask'].float().sum(-1) + 1).unsqueeze(-1))
            ent = (ent * su_mask).sum(-1) / (su_mask.sum(-1) + 1e-9)
        elif h == 'target_unit':
            ent = ent / (1e-9 + torch.log(mask['target_units_logits_mask'].float().sum(-1) + 1))
        else:
            ent = ent / math.log(logits[h].shape[-1])
        ent = (ent * am(h)).mean()
        info['entropy/' + h] = ent.detach()
        total_ent = total_ent - ent * cfg['head_w'][h]
    info['entropy/total'] = total_ent.detach()
    total_ent = total_ent * cfg['entropy_w']
    total_kl = 0.                                                        # as_rl_utils.py:75-103
    at_kl = None
    for h in HEADS:
        tlp = torch.log_softmax(teacher[h], dim=-1)
        kl = (tlp.exp() * (tlp - logp_all[h])).sum(-1)
        if h == 'selected_units':
            kl = (kl * su_mask).sum(-1)
        kl = kl * am(h)
        if h == 'action_type':
            at_kl = (kl * (step < cfg['action_type_kl_steps']) * mask['cum_action_mask']).mean()
            info['kl/extra_at'] = at_kl.detach()
        kl = kl.mean()
        info['kl/' + h] = kl.detach()
        total_kl = total_kl + kl * cfg['head_w'][h]
    info['kl/total'] = total_kl.detach()
    total_kl = total_kl * cfg['kl_w']
    at_kl = at_kl * cfg['action_type_kl_w']
    total_dapo = 0.
    if use_dapo:                                                         # as_rl_utils.py:105-127
        early = step < dapo_steps
        for h in HEADS:
            slp = torch.log_softmax(out['successive_logit'][h], dim=-1)
            kl = (slp.exp() * (slp - logp_all[h])).sum(-1)
            if h == 'selected_units':
                kl = (kl * su_mask).sum(-1)
            kl = (kl * am(h) * early).mean()
            info['battle/' + h] = kl.detach()                            # logged under 'battle/*' (overrides the pg entries)
            total_dapo = total_dapo + kl * cfg['head_w'][h]
        info['battle/total'] = total_dapo.detach()
        total_dapo = total_dapo * dapo_w
    if only_update_value:
        info['total_loss'] = total_critic
    else:
        info['total_loss'] = total_pg + total_upgo + total_critic + total_ent + total_kl + at_kl + total_dapo
    return info


# --------------------------------------------------------------------------------------------
# SL loss (sl_training/sl_loss.py:100-286, su_mask False as in bin/sl_user_config.yaml:28)
# --------------------------------------------------------------------------------------------
SL_LOSS_WEIGHTS = {'action_type': 30.0, 'delay': 9.0, 'queued': 1.0, 'selected_units': 4.0, 'target_unit': 4.0,
                   'target_location': 8.0}


def sl_loss(logits: dict, actions: dict, actions_mask: dict, selected_units_num: Tensor, entity_num: Tensor = None,
            infer_selected_units: Tensor = None, su_mask: bool = False, label_smooth: bool = False) -> Dict[str, Tensor]:
    """SupervisedLoss.compute_loss, sl_loss.py:100-286: the six losses plus the no-grad metrics."""
    out = {}

    def criterion(x, y):                                                 # sl_loss.py:16-34,54-57
        if not label_smooth:
            return F.cross_entropy(x, y, reduction='none')
        lp = torch.log_softmax(x, dim=-1)
        return 0.9 * (-lp.gather(-1, y.unsqueeze(1)).squeeze(1)) + 0.1 * (-lp.mean(dim=-1))

    for h in HEADS:
        mask = actions_mask[h]
        if h == 'selected_units':                                        # sl_loss.py:174-238
            lg, labels = logits[h], actions[h]
            b, s, n = lg.shape
            if su_mask:                                                  # :177-192: other selected units masked at every step
                keep = sequence_mask((selected_units_num - 1).clamp(min=0), labels.shape[1])
                nl = labels.clone()
                nl[~keep] = n
                nl = nl[:, :s]
                lg = torch.cat([lg, torch.zeros(b, s, 1)], dim=-1)
                lm = torch.ones_like(lg)
                lm = torch.scatter(lm, 2, nl.unsqueeze(1).repeat(1, s, 1), 0.)
                lm = torch.scatter(lm, 2, nl.unsqueeze(2), 1.)
                lg = lg.masked_fill(~lm.bool(), -1e9)[:, :, :-1]
            select = sequence_mask(selected_units_num, s)
            ce = F.cross_entropy(lg.reshape(-1, n), labels[:, :s].reshape(-1), reduction='none').view(b, s)
            ce = ce.masked_fill(~select, 0) * mask.unsqueeze(1)
            out[h + '_loss'] = ce.sum() / b
            out['selected_units_loss_norm'] = (ce.sum() / (selected_units_num.sum() + 1e-6)).detach()
            out['selected_units_end_flag_loss'] = ce[torch.arange(b), selected_units_num - 1].mean().detach()
            if infer_selected_units is not None:                         # :206-232
                preds = infer_selected_units
                end = (preds == entity_num.unsqueeze(1)).long().argmax(dim=-1)
                invalid = end == 0
                end = end + 1
                end[invalid] += s
                pm = sequence_mask(end, s)
                lab = (labels[:, :s] + 1) * select
                prd = (preds + 1) * pm
                ps = torch.zeros(b, n + 1, dtype=torch.bool).scatter(1, prd.long(), True)
                ls = torch.zeros(b, n + 1, dtype=torch.bool).scatter(1, lab.long(), True)
                inter, union = (ps & ls)[:, 1:].sum(1), (ps | ls)[:, 1:].sum(1)
                out['selected_units_iou'] = (inter / (union + 1e-6) * mask).sum() / (mask.sum() + 1e-6)
            else:
                out['selected_units_iou'] = torch.tensor(0.)
            continue
        ce = criterion(logits[h], actions[h]) * mask
        valid = mask.sum()
        out[h + '_loss'] = ce.sum() / valid if valid > 0 else ce.sum() * 0
        with torch.no_grad():
            pred = logits[h].argmax(dim=-1)
            if h == 'action_type':
                out['action_type_acc'] = (pred == actions[h]).float().sum() / len(actions[h])
            elif h == 'delay':
                out['delay_distance_L1'] = ((pred - actions[h]).abs() * mask).sum() / (mask.sum() + 1e-6)
            elif h == 'queued':
                out['queued_acc'] = ((pred - actions[h]).abs() * mask).sum() / (mask.sum() + 1e-6)
            elif h == 'target_unit':
                out['target_unit_acc'] = ((pred == actions[h]) * mask).sum() / (mask.sum() + 1e-6)
            else:
                W = 160                                                  # hard-coded in the reference (sl_loss.py:257)
                d = ((pred % W - actions[h] % W) ** 2 + (pred // W - actions[h] // W) ** 2).float().sqrt()
                out['target_location_distance_L2'] = (d * mask).sum() / (mask.sum() + 1e-6)
    out['total_loss'] = sum(out[h + '_loss'] * SL_LOSS_WEIGHTS[h] for h in HEADS)
    return out
