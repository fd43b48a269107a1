"""Host-side dataflow of the AlphaStar policy: encoders -> core LSTM -> auto-regressive heads -> baselines.

This file is the *orchestration* of the hot path: it decides shapes, masks and the
order of work, and hands the heavy steps to the kernels in ``ops`` (tcgen05 split GEMM for every tileable
fc_block, scatter_connection, categorical sampling ...).  Steps that have no hand-written kernel yet are
plain device-side torch calls (cuBLAS / cuDNN / ATen: library code, listed per row in DESIGN.md §coverage).
Parameters arrive as ``P[name]`` with the reference's state_dict names.

Semantics follow the reference (DI-star ``distar/agent/default/model``) including its quirks:
  * pooled entity mean uses relu(x) because the reference's shared ReLU is in-place (entity_encoder.py:39,81-85)
  * effect planes always light flat pixel 0 (zero-padded index lists, spatial_encoder.py:62-69)
  * (x, y) clamped in the scatter (module_utils.py:18-19); learned end token at slot entity_num
    (action_arg_head.py:118-129); LSTM carries the layer-normed cell state (lstm.py:150)
  * teacher-forced selected-units loop runs max(selected_units_num) steps for the whole batch, rows with
    selected_units_num == 0 are not normalised (action_arg_head.py:179-200); logits padded to 64 with -1e9
    (model.py:156-158); delay logits are not divided by the temperature (action_arg_head.py:41-47)
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor
HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']
MAX_SELECTED_UNITS_NUM = 64
MAX_ENTITY_NUM = 512
NUM_ACTIONS = 327

ENTITY_FIELDS = [  # (name, kind, width): actor_critic_default_config.yaml:264-364
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
SCALAR_FIELDS = [  # (name, kind, vocab/in, ctx, baseline): yaml:146-219, scalar_encoder.py:99-132
    ('agent_statistics', 'fc', 10, False, True), ('home_race', 'emb', 5, True, False),
    ('away_race', 'emb', 5, True, False), ('upgrades', 'fc', 90, False, True),
    ('unit_counts_bow', 'fc', 260, False, True), ('last_delay', 'emb', 128, False, False),
    ('last_queued', 'emb', 2, False, False), ('last_action_type', 'emb', 327, False, False),
    ('cumulative_stat', 'fc', 167, True, True), ('beginning_order', 'bo', 214, True, True),
    ('unit_type_bool', 'fc', 260, True, False), ('enemy_unit_type_bool', 'fc', 260, True, False),
    ('unit_order_type', 'fc', 269, True, False),
]
SPATIAL_ONEHOT = [('visibility_map', 4), ('creep', 2), ('player_relative', 5), ('alerts', 2), ('pathable', 2),
                  ('buildable', 2)]
SPATIAL_EFFECTS = ['effect_PsiStorm', 'effect_NukeDot', 'effect_LiberatorDefenderZone', 'effect_BlindingCloud',
                   'effect_CorrosiveBile', 'effect_LurkerSpines']
BASELINE_ATAN = {'winloss': True, 'build_order': False, 'built_unit': False, 'effect': False, 'upgrade': False,
                 'battle': False}


VALUE_FC_FIELDS = ['enemy_unit_counts_bow', 'enemy_unit_type_bool', 'enemy_agent_statistics', 'enemy_upgrades', 'cumulative_stat']


def set_library_precision():
    """The reference computes in fp32 end to end.  PyTorch lets cuDNN convolutions silently use TF32 (10-bit mantissa),
    which alone costs ~1e-3 on the location logits; the parts of the path that are still library calls must run in
    true fp32 to hold the 1e-3 parity contract."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def bad_input_message(code: int) -> str:
    """device error flag -> the message of the exception the reference raises at that point"""
    what = []
    if code & 1:
        what.append('negative categorical id in an entity field')                    # entity_encoder.py:69-72
    if code & 4:
        what.append('action / scalar id outside its vocabulary (one-hot index out of range)')   # F.one_hot / nn.Embedding raise
    return '; '.join(what) or 'invalid input (code %d)' % code


class Net:
    """Functional network over a parameter mapping.  ``terms`` = products per tensor-core GEMM (3 = fp32-class)."""

    def __init__(self, P: Dict[str, Tensor], spatial_x: int, spatial_y: int, temperature: float = 1.0,
                 terms: int = 3, rng: str = 'cuda'):
        self.P, self.W, self.H, self.T, self.terms, self.rng = P, spatial_x, spatial_y, temperature, terms, rng
        self.bad_input_flag: Optional[Tensor] = None      # device int32[1]: set by kernels that meet an invalid input
        set_library_precision()

    def raise_on_bad_input(self) -> None:
        """entity_encoder.py:69-72 raises on a negative categorical id.  The kernels only record it (one device flag for
        the whole forward) so the check costs a single host read at the end instead of one pipeline drain per chunk."""
        if self.bad_input_flag is not None:
            code = int(self.bad_input_flag.item())
            if code != 0:
                self.bad_input_flag.zero_()
                raise RuntimeError(bad_input_message(code))

    # -------------------------------------------------------------------------------------- primitives
    def fc(self, name: str, x: Tensor, relu: bool = False, split: bool = False, exact: bool = False) -> Tensor:
        """fc_block; split=True makes the GEMM epilogue also write the bf16 pair its consumer GEMM will read, split='only'
        writes nothing but the pair (the consumer must be another tcgen05 GEMM; the fp32 result is a NaN placeholder).
        Shapes the tile grid does not divide (and integer-typed observation inputs) go through the zero-padded tensor-core
        path (ops.linear_any); exact=True: the input values are exact in bf16 (0/1 flags, counts <= 256)."""
        return ops.linear(x, self.P[name + '.0.weight'], self.P[name + '.0.bias'], relu, self.terms, split, exact_input=exact)

    def conv(self, name: str, x: Tensor, pad: int, relu: bool = False) -> Tensor:
        y = F.conv2d(x, self.P[name + '.0.weight'], self.P[name + '.0.bias'], padding=pad)
        return torch.relu(y) if relu else y

    def ln(self, name: str, x: Tensor, residual: Optional[Tensor] = None, split: bool = False) -> Tensor:
        """LayerNorm(x [+ residual]); split=True also emits the bf16 pair for the GEMM that consumes the result."""
        return ops.layer_norm(x, self.P[name + '.weight'], self.P[name + '.bias'], residual, split)

    def sample(self, logits: Tensor) -> Tensor:
        return ops.sample_categorical(logits, rng=self.rng)[0]

    # -------------------------------------------------------------------------------------- transformer
    def attention(self, pre: str, x: Tensor, key_mask: Optional[Tensor], heads: int, hd: int) -> Tensor:
        """module_utils.py:88-111."""
        B, N, _ = x.shape
        if key_mask is None and x.is_cuda:          # the 20-token beginning-build-order transformer: one warp per (row, head)
            return self.fc(pre + '.project', ops.small_attention(self.fc(pre + '.attention_pre', x), heads, hd))
        q, k, v = self.fc(pre + '.attention_pre', x).view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        score = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
        if key_mask is not None:
            score = score.masked_fill(~key_mask.view(B, 1, 1, N), -1e9)
        a = torch.matmul(torch.softmax(score, dim=-1), v).permute(0, 2, 1, 3).reshape(B, N, heads * hd)
        return self.fc(pre + '.project', a)

    def transformer(self, pre: str, x: Tensor, key_mask, heads: int, hd: int, post_ln: bool, embedded: bool = False) -> Tensor:
        """module_utils.py:130-151,191-199.  embedded: x already went through the embedding layer."""
        if not embedded:
            x = self.fc(pre + '.embedding', x, relu=True)
        for i in range(3):
            lp = '%s.layers.%d' % (pre, i)
            if post_ln:
                x = self.ln(lp + '.layernorm1', x + self.attention(lp + '.attention', x, key_mask, heads, hd))
                m = self.fc(lp + '.mlp.1', self.fc(lp + '.mlp.0', x, relu=True), relu=True)
                x = self.ln(lp + '.layernorm2', x + m)
            else:
                x = x + self.attention(lp + '.attention', self.ln(lp + '.layernorm1', x), key_mask, heads, hd)
                x = x + self.fc(lp + '.mlp.1', self.fc(lp + '.mlp.0', self.ln(lp + '.layernorm2', x), relu=True),
                                relu=True)
        return x

    # -------------------------------------------------------------------------------------- encoders
    def bo_embedding(self, pre: str, bo: Tensor, loc: Tensor) -> Tensor:
        """BeginningBuildOrderEncoder.forward (scalar_encoder.py:34-53); `pre` selects the scalar encoder's or the value
        encoder's copy of the module."""
        P = self.P
        if bo.is_cuda:
            # token features straight into the embedding GEMM's exact bf16 operand (all of them are 0 / 1)
            tp = pre + 'transformer'
            hi = ops.bo_tokens(bo, loc, self.W)
            w = P[tp + '.embedding.0.weight']
            x0 = ops.linear_presplit(hi, None, F.pad(w, (0, hi.shape[1] - w.shape[1])), P[tp + '.embedding.0.bias'], True,
                                     self.terms).view(bo.shape[0], bo.shape[1], -1)
            t = self.transformer(tp, x0, None, 2, 8, post_ln=False, embedded=True)
            return self.fc(pre + 'embedd_fc', t.mean(dim=1), relu=True)
        bo, loc = bo.long(), loc.long()
        B, dev = bo.shape[0], bo.device
        bits = torch.arange(9, -1, -1, device=dev)
        tok = torch.cat([F.one_hot(bo, 174).float(),
                         torch.eye(20, device=dev).unsqueeze(0).expand(B, -1, -1),
                         (((loc % self.W).unsqueeze(-1) >> bits) & 1).float(),
                         (((loc // self.W).unsqueeze(-1) >> bits) & 1).float()], dim=2)
        t = self.transformer(pre + 'transformer', tok, None, 2, 8, post_ln=False)
        return self.fc(pre + 'embedd_fc', t.mean(dim=1), relu=True)

    def scalar_encoder(self, s: Dict[str, Tensor]):
        """obs_encoder/scalar_encoder.py:99-132 (K8)."""
        P, pre = self.P, 'encoder.scalar_encoder.encode_modules.'
        dev = s['time'].device
        outs, ctx, base = [], [], []
        for name, kind, din, is_ctx, is_base in SCALAR_FIELDS:
            if kind == 'emb':       # relu(nn.Embedding(clamp(id))) as one gather kernel (scalar_encoder.py:105-116)
                e = ops.onehot_linear(P[pre + name + '.weight'], None, s[name], relu=True, embedding=True, clamp_max=True,
                                      flag=self.bad_input_flag)
            elif kind == 'fc':      # the observation goes to the GEMM in its wire dtype: 0/1 flags and counts are exact in bf16
                x = s[name]
                e = self.fc(pre + name, x if x.is_cuda else x.float(), relu=True,
                            exact=x.dtype in (torch.uint8, torch.int8, torch.int16))
            else:
                e = self.bo_embedding(pre + 'beginning_order.', s['beginning_order'], s['bo_location'])
            outs.append(e)
            if is_ctx:
                ctx.append(e)
            if is_base:
                base.append(e)
        pa = P['encoder.scalar_encoder.position_array']
        t = s['time'].float().unsqueeze(1)
        te = torch.stack([torch.sin(t * pa[0::2]), torch.cos(t * pa[1::2])], dim=2).reshape(t.shape[0], -1)
        outs.append(te)
        return torch.cat(outs, 1), torch.cat(ctx, 1), torch.cat(base, 1)

    def entity_features(self, e: Dict[str, Tensor], pad_to: int = 1024) -> Tensor:
        """997-wide expansion of entity_encoder.py:59-78, zero padded to a GEMM-tileable width."""
        dev = e['x'].device
        cols = []
        bits = torch.arange(10, -1, -1, device=dev)
        for name, kind, w in ENTITY_FIELDS:
            v = e[name]
            if kind == 'o':
                cols.append(F.one_hot(v.long().clamp(max=w - 1), w).float())
            elif kind == 'b':
                cols.append(((v.long().unsqueeze(-1) >> bits) & 1).float())
            else:
                cols.append(v.float().unsqueeze(-1))
        width = sum(c.shape[-1] for c in cols)
        if pad_to > width:
            cols.append(torch.zeros(*cols[0].shape[:-1], pad_to - width, device=dev))
        return torch.cat(cols, dim=-1)

    def entity_encoder(self, e: Dict[str, Tensor], entity_num: Tensor):
        """obs_encoder/entity_encoder.py:59-96 (K1-K4)."""
        P, pre = self.P, 'encoder.entity_encoder.'
        w = P[pre + 'transformer.embedding.0.weight']
        E = entity_info_E = e['x'].shape[1]
        mask = torch.arange(E, device=e['x'].device).unsqueeze(0) < entity_num.unsqueeze(1)
        if self.bad_input_flag is None and e['x'].is_cuda:
            self.bad_input_flag = torch.zeros(1, dtype=torch.int32, device=e['x'].device)
        split = ops.entity_features_split(e, ENTITY_FIELDS, flag=self.bad_input_flag, exact=True)
        if split is not None:       # K1: features expanded straight into the GEMM's (exact) bf16 operand, no lo half
            w_x = ops.entity_exact_weight(w, ENTITY_FIELDS)
            x = ops.linear_presplit(split[0], None, w_x, P[pre + 'transformer.embedding.0.bias'], True, self.terms,
                                    emit_split=True)
        else:
            for name, kind, wd in ENTITY_FIELDS:
                if kind == 'o' and e[name].dtype in (torch.int8, torch.int16, torch.int32, torch.int64):
                    if bool((e[name] < 0).any()):
                        raise RuntimeError('negative categorical id in entity field %s' % name)
            x = ops.linear(self.entity_features(e), F.pad(w, (0, 1024 - w.shape[1])), P[pre + 'transformer.embedding.0.bias'],
                           True, self.terms)
        for i in range(3):
            lp = '%stransformer.layers.%d' % (pre, i)
            # the two consumers of x in each sub-layer (the GEMM and the residual input of the LayerNorm) are routed through
            # the GEMM's autograd node (fork=True): its dX epilogue adds the residual branch's gradient
            qkv, xr = ops.linear(x, P[lp + '.attention.attention_pre.0.weight'], P[lp + '.attention.attention_pre.0.bias'], False,
                                 self.terms, 'only' if x.is_cuda else False, fork=True)
            a = self.fc(lp + '.attention.project', ops.entity_attention(qkv, entity_num, 2, 128,
                                                                       P[lp + '.attention.attention_pre.0.bias']))
            x = self.ln(lp + '.layernorm1', xr, residual=a, split=True)
            m, xr = ops.ffn(x, P[lp + '.mlp.0.0.weight'], P[lp + '.mlp.0.0.bias'], P[lp + '.mlp.1.0.weight'],
                            P[lp + '.mlp.1.0.bias'], self.terms, fork=True)
            x = self.ln(lp + '.layernorm2', xr, residual=m, split=(i < 2))
        x = torch.relu(x)
        entity_embeddings = self.fc(pre + 'entity_fc', x, relu=True)
        # masked mean over entities as a batched [1,E] x [E,256] product: reads x once instead of materialising x * mask
        pooled = torch.bmm(mask.to(x.dtype).unsqueeze(1), x).squeeze(1) / entity_num.unsqueeze(-1)
        return entity_embeddings, self.fc(pre + 'embed_fc', pooled, relu=True), mask

    def spatial_encoder(self, sp: Dict[str, Tensor], project: Tensor, ex: Tensor, ey: Tensor, entity_num: Tensor):
        """obs_encoder/spatial_encoder.py:51-90 (K7) with the entity scatter (K6) fused into its first stage."""
        pre = 'encoder.spatial_encoder.'
        N = project.shape[0]
        # stem: scatter + plane expansion + 1x1 project conv + ReLU + first 2x2 max-pool in one kernel; from here on
        # channels-last, channels padded to 64, every 3x3 conv an implicit GEMM on the tensor cores
        x = ops.spatial_stem(sp, project, ex, ey, entity_num, self.P[pre + 'project.0.weight'],
                             self.P[pre + 'project.0.bias'], 64)
        skips = [None, None, None]            # the 128^2 / 64^2 / 32^2 skips are never read downstream (unet off)
        for i in range(3):
            if i > 0:
                x = self.pool_nhwc(x)
            x = self.conv_nhwc(pre + 'downsample.%d' % i, x, relu=True)
        for i in range(4):
            skips.append(x)
            r = self.conv_nhwc(pre + 'res.%d.conv1' % i, x, relu=True, split='only')
            x = self.conv_nhwc(pre + 'res.%d.conv2' % i, r, relu=True, residual=x, split=(i < 3))   # relu(conv2(r) + x)
        h8, w8, c = x.shape[1:]
        w = self.P[pre + 'fc.0.weight']
        w = w.view(w.shape[0], c, h8, w8).permute(0, 2, 3, 1).reshape(w.shape[0], -1)    # (c,y,x) -> (y,x,c) columns
        return ops.linear(x.reshape(N, -1), w, self.P[pre + 'fc.0.bias'], True, self.terms), skips

    def conv_nhwc(self, name: str, x: Tensor, relu: bool = False, residual: Optional[Tensor] = None,
                  split: bool = False) -> Tensor:
        return ops.conv_nhwc(x, self.P[name + '.0.weight'], self.P[name + '.0.bias'], relu, residual, self.terms, split)

    @staticmethod
    def pool_nhwc(x: Tensor) -> Tensor:
        return ops.max_pool2_nhwc(x)

    def encoder(self, spatial_info, entity_info, scalar_info, entity_num, entity_fn=None, scalar_out=None):
        """model/encoder.py:28-45.  entity_fn lets the caller wrap the entity transformer (the activation-memory hog)
        in activation checkpointing while the rest of the encoder keeps its activations.  scalar_out: this chunk's rows of a
        scalar encoder pass the caller already ran over ALL observation rows (its ~20 small layers are launch-bound: once
        per step instead of once per encoder chunk)."""
        embedded_scalar, scalar_context, baseline_feature = scalar_out if scalar_out is not None else \
            self.scalar_encoder(scalar_info)
        run_entity = entity_fn or self.entity_encoder
        entity_embeddings, embedded_entity, _mask = run_entity(entity_info, entity_num)
        if entity_embeddings.is_cuda:
            # 256 -> 32: too narrow for a tensor-core tile, so the weight is zero-padded to 64 rows (one 64-wide tcgen05 GEMM each
            # for forward, dX and dW instead of three skinny fp32 library GEMMs) and the real 32 columns are sliced back out
            P = self.P
            w = F.pad(P['encoder.scatter_project.0.weight'], (0, 0, 0, 32))
            b = F.pad(P['encoder.scatter_project.0.bias'], (0, 32))
            project = ops.linear(entity_embeddings, w, b, True, self.terms, allow_n64=True)[..., :32].contiguous()
        else:
            project = self.fc('encoder.scatter_project', entity_embeddings, relu=True)
        # scatter_connection (ops.scatter_connection, K6) is fused into the spatial stem; the stand-alone operator is
        # kept for API parity / measurement
        embedded_spatial, map_skip = self.spatial_encoder(spatial_info, project, entity_info['x'], entity_info['y'],
                                                          entity_num)
        lstm_input = torch.cat([embedded_scalar, embedded_entity, embedded_spatial], dim=-1)
        return lstm_input, scalar_context, baseline_feature, entity_embeddings, map_skip

    # -------------------------------------------------------------------------------------- LSTM
    def lstm_cell(self, pre: str, ig: Tensor, h: Tensor, c: Tensor):
        """LayerNormLSTMCell with the input half (LN_i(x W_ih^T)) precomputed: lstm.py:138-153."""
        P = self.P
        return ops.lstm_cell(ig, h @ P[pre + '.weight_hh'].t(), c, P[pre + '.layernorm_h.weight'],
                             P[pre + '.layernorm_h.bias'], P[pre + '.layernorm_c.weight'], P[pre + '.layernorm_c.bias'])

    def lstm(self, pre: str, x: Tensor, state: List[Tuple[Tensor, Tensor]], layers: int):
        """StackedLSTM over [L,B,D] (lstm.py:161-167,223-234).  The input projection of a whole layer is one
        GEMM over all timesteps (the layer-major loop order of the reference makes that legal)."""
        out_state = []
        for l in range(layers):
            cp = '%s.layers.%d.cell' % (pre, l)
            L, B, D = x.shape
            ig = self.ln(cp + '.layernorm_i', ops.linear(x.reshape(L * B, D), self.P[cp + '.weight_ih'], None, False,
                                                         self.terms)).view(L, B, -1)
            h, c = state[l]
            if ig.is_cuda and h.shape[-1] in (128, 384):
                P = self.P
                x, c = ops.lstm_layer(ig, h, c, P[cp + '.weight_hh'], P[cp + '.layernorm_h.weight'], P[cp + '.layernorm_h.bias'],
                                      P[cp + '.layernorm_c.weight'], P[cp + '.layernorm_c.bias'])
                h = x[L - 1]
            else:
                ys = []
                for t in range(L):
                    h, c = self.lstm_cell(cp, ig[t], h, c)
                    ys.append(h)
                x = torch.stack(ys)
            out_state.append((h, c))
        return x, out_state

    # -------------------------------------------------------------------------------------- heads
    def glu(self, name: str, x: Tensor, ctx: Tensor) -> Tensor:
        """GLU (module_utils.py:508-524): layer2(sigmoid(layer1(context)) * x)."""
        return self.fc(name + '.layer2', ops.glu_gate(self.fc(name + '.layer1', ctx), x))

    def action_type_head(self, lstm_out, scalar_context, action_type=None):
        """head/action_type_head.py:48-67 (K10)."""
        pre = 'policy.action_type_head.'
        x = self.fc(pre + 'project', lstm_out, relu=True)
        for i in range(2):
            r = torch.relu(self.ln(pre + 'res.%d.fc1.1' % i, self.fc(pre + 'res.%d.fc1' % i, x)))
            r = self.ln(pre + 'res.%d.fc2.1' % i, self.fc(pre + 'res.%d.fc2' % i, r))
            x = torch.relu(r + x)
        logits = self.glu(pre + 'action_fc', x, scalar_context) / self.T
        if action_type is None:
            action_type = self.sample(logits)
        # relu(fc(one_hot(a))) == gather of weight columns (+ bias, ReLU) in one kernel; an id outside the head is recorded
        e1 = ops.onehot_linear(self.P[pre + 'action_map_fc1.0.weight'], self.P[pre + 'action_map_fc1.0.bias'], action_type,
                               relu=True, flag=self.bad_input_flag)
        e1 = self.glu(pre + 'glu1', self.fc(pre + 'action_map_fc2', e1), scalar_context)
        return logits, action_type, e1 + self.glu(pre + 'glu2', lstm_out, scalar_context)

    def arg_head(self, pre: str, emb, n: int, use_temperature: bool, action=None):
        """DelayHead / QueuedHead: head/action_arg_head.py:41-53,73-86 (K11)."""
        x = self.fc(pre + 'fc3', self.fc(pre + 'fc2', self.fc(pre + 'fc1', emb, relu=True), relu=True))
        if use_temperature:
            x = x / self.T
        if action is None:
            action = self.sample(x)
        e = ops.onehot_linear(self.P[pre + 'embed_fc1.0.weight'], self.P[pre + 'embed_fc1.0.bias'], action, relu=True,
                              flag=self.bad_input_flag)
        return x, action, emb + self.fc(pre + 'embed_fc2', e)

    def head_keys(self, entity_embeddings):
        """key_fc of the selected-units head and of the target-unit head (action_arg_head.py:118-129, 343-349): two 256 -> 32
        projections of the same [P, 512, 256] entity embeddings.  32 outputs are too narrow for a tensor-core tile and ran as
        skinny fp32 library GEMMs (6.6 ms / step with their gradients); stacked they are one 64-wide tcgen05 GEMM."""
        cached = getattr(self, '_head_keys', None)
        if cached is not None and cached[0] is entity_embeddings:
            return cached[1], cached[2]
        P = self.P
        a, b = 'policy.selected_units_head.key_fc.0.', 'policy.target_unit_head.key_fc.0.'
        kfull = None
        if entity_embeddings.is_cuda:
            w = torch.cat([P[a + 'weight'], P[b + 'weight']], dim=0)
            bias = torch.cat([P[a + 'bias'], P[b + 'bias']], dim=0)
            kfull = ops.linear(entity_embeddings, w, bias, False, self.terms, allow_n64=True)
            # the three consumers of the stacked keys accumulate their gradients into one shared buffer (ops.KeyGradSink)
            kfull, self._head_keys_sink = ops.fork_keys(kfull)
            ksu, ktu = kfull[..., :32], kfull[..., 32:]
        else:
            ksu = self.fc('policy.selected_units_head.key_fc', entity_embeddings)
            ktu = self.fc('policy.target_unit_head.key_fc', entity_embeddings)
        self._head_keys = (entity_embeddings, ksu, ktu)
        self._head_keys_full = kfull
        if kfull is None:
            self._head_keys_sink = None
        return ksu, ktu

    def su_keys(self, entity_embeddings, entity_num):
        """_get_key_mask, action_arg_head.py:118-143."""
        pre = 'policy.selected_units_head.'
        N, E, _ = entity_embeddings.shape
        key = self.head_keys(entity_embeddings)[0]
        slot = torch.arange(E + 1, device=key.device).unsqueeze(0)
        is_end = (slot == entity_num.unsqueeze(1)).unsqueeze(-1)
        key = torch.where(is_end, self.P[pre + 'end_embedding'].view(1, 1, -1),
                          F.pad(key, (0, 0, 0, 1)))
        return key, slot < (entity_num + 1).unsqueeze(1), slot

    def su_embed(self, key, weights, normalise):
        """masked mean of selected keys -> embed_fc2(relu(embed_fc1)): action_arg_head.py:196-199,290-293.
        weights [..., E+1] (0/1), key [N,E+1,32]."""
        pre = 'policy.selected_units_head.'
        s = torch.matmul(weights, key) if weights.dim() == 3 else (key * weights.unsqueeze(2)).sum(dim=1)
        cnt = weights.sum(dim=-1, keepdim=True)
        s = torch.where(normalise, s / cnt, s)
        return self.fc(pre + 'embed_fc2', self.fc(pre + 'embed_fc1', s, relu=True))

    def selected_units_train(self, emb0, entity_embeddings, entity_num, selected_units_num, selected_units,
                             steps: Optional[int] = None):
        """Teacher-forced pointer network (action_arg_head.py:168-216) in its step-parallel form
        (SURVEY.md Appendix B.1): everything but the 32-wide LN-LSTM is computed for all steps at once.

        The reference loops max(selected_units_num) times.  ``steps`` is that maximum when the caller already has it on the
        host (the learner fetches it with an asynchronous copy queued before the encoder, so reading it here does not
        drain the launch queue); otherwise it is read back now."""
        pre = 'policy.selected_units_head.'
        N = emb0.shape[0]
        S = max(int(selected_units_num.max()) if steps is None else int(steps), 1)
        if emb0.is_cuda:
            # K12, training path: the three non-GEMM pieces are one kernel each (csrc/su_train.cu), the four small MLP layers
            # run on the tensor cores over all (row, step) pairs; keys are read in place from the stacked key projection
            P_, cp = self.P, pre + 'lstm.layers.0.cell'
            self.head_keys(entity_embeddings)
            kfull, sink = self._head_keys_full, self._head_keys_sink
            su = selected_units.long().contiguous()
            en, num = entity_num.to(torch.int64).contiguous(), selected_units_num.to(torch.int64).contiguous()
            mean = ops.su_prefix_mean(kfull, su, en, num, S, sink)                                 # [N,S,32]
            emb_steps = emb0.unsqueeze(1) + self.fc(pre + 'embed_fc2', self.fc(pre + 'embed_fc1', mean, relu=True))
            ae = torch.cat([emb0.unsqueeze(1), emb_steps[:, :-1]], dim=1)
            q = self.fc(pre + 'query_fc2', self.fc(pre + 'query_fc1', ae, relu=True))              # [N,S,32]
            ig = self.ln(cp + '.layernorm_i', ops.linear(q, P_[cp + '.weight_ih'], None, False, self.terms))
            hs = ops.su_lstm(ig, P_[cp + '.weight_hh'], P_[cp + '.layernorm_h.weight'], P_[cp + '.layernorm_h.bias'],
                             P_[cp + '.layernorm_c.weight'], P_[cp + '.layernorm_c.bias'])
            logits = ops.su_logits(hs, kfull, P_[pre + 'end_embedding'], su, en, sink)
            return logits, emb_steps[:, -1], selected_units_num
        key, valid, slot = self.su_keys(entity_embeddings, entity_num)
        su = selected_units[:, :S].long()
        onehot = su.unsqueeze(-1) == slot.unsqueeze(1)                               # [N,S,E+1]
        ended = torch.cummax((su == entity_num.unsqueeze(1)).long(), dim=1)[0].bool()  # end_flag after step i
        picked = torch.cummax((onehot & ~ended.unsqueeze(-1)).long(), dim=1)[0].float()  # cumulative one-hot
        normalise = (selected_units_num != 0).view(N, 1, 1)
        emb_steps = emb0.unsqueeze(1) + self.su_embed(key, picked, normalise)         # ae after step i
        ae = torch.cat([emb0.unsqueeze(1), emb_steps[:, :-1]], dim=1)                 # ae feeding step i
        q = self.fc(pre + 'query_fc2', self.fc(pre + 'query_fc1', ae, relu=True))     # [N,S,32]
        cp = pre + 'lstm.layers.0.cell'
        ig = self.ln(cp + '.layernorm_i', q @ self.P[cp + '.weight_ih'].t())
        h = torch.zeros(N, 32, device=emb0.device)
        c = torch.zeros(N, 32, device=emb0.device)
        hs = []
        for i in range(S):
            h, c = self.lstm_cell(cp, ig[:, i], h, c)
            hs.append(h)
        logits = torch.matmul(torch.stack(hs, dim=1), key.transpose(1, 2))            # [N,S,E+1]
        # mask recurrence: step 0 end slot off; step i>=1: all valid slots minus every unit chosen before i
        chosen_before = torch.cummax(onehot.long(), dim=1)[0].bool()
        chosen_before = torch.cat([torch.zeros_like(chosen_before[:, :1]), chosen_before[:, :-1]], dim=1)
        step_mask = valid.unsqueeze(1) & ~chosen_before
        step_mask[:, 0] = valid & (slot != entity_num.unsqueeze(1))
        logits = logits.masked_fill(~step_mask, -1e9)
        return logits, emb_steps[:, -1], selected_units_num

    def selected_units_sample(self, emb0, entity_embeddings, entity_num, su_mask, fixed_steps: Optional[int] = None):
        """Sampling pointer network, action_arg_head.py:262-314 (K12; sequential, early exit when all rows ended)."""
        pre = 'policy.selected_units_head.'
        N = emb0.shape[0]
        dev = emb0.device
        rows = torch.arange(N, device=dev)
        key, valid, slot = self.su_keys(entity_embeddings, entity_num)
        if emb0.is_cuda:
            # K12: one fused kernel per step (query MLP + LN-LSTM + dot/mask/sample + bookkeeping + embedding MLP)
            P, cp = self.P, pre + 'lstm.layers.0.cell'
            w16 = [P[pre + 'query_fc1.0.weight'], P[pre + 'query_fc1.0.bias'], P[pre + 'query_fc2.0.weight'],
                   P[pre + 'query_fc2.0.bias'], P[cp + '.weight_ih'], P[cp + '.weight_hh'],
                   P[cp + '.layernorm_i.weight'], P[cp + '.layernorm_i.bias'], P[cp + '.layernorm_h.weight'],
                   P[cp + '.layernorm_h.bias'], P[cp + '.layernorm_c.weight'], P[cp + '.layernorm_c.bias'],
                   P[pre + 'embed_fc1.0.weight'], P[pre + 'embed_fc1.0.bias'], P[pre + 'embed_fc2.0.weight'],
                   P[pre + 'embed_fc2.0.bias']]
            if fixed_steps:          # graph-capturable form: a fixed number of pointer steps, no host reads
                logits, units, ae, num = ops.su_sample(w16, emb0, key, valid, entity_num, su_mask, self.T, self.rng,
                                                       max_steps=fixed_steps, poll=False)
            else:
                logits, units, ae, num = ops.su_sample(w16, emb0, key, valid, entity_num, su_mask, self.T, self.rng)
            return logits, units, ae, num, torch.zeros(N, MAX_ENTITY_NUM + 1, device=dev)
        step_mask = valid & (slot != entity_num.unsqueeze(1))
        num = torch.full((N,), MAX_SELECTED_UNITS_NUM, dtype=torch.long, device=dev)
        num[~su_mask] = 0
        end_flag = ~su_mask
        picked = torch.zeros(N, key.shape[1], device=dev)
        cp = pre + 'lstm.layers.0.cell'
        h = torch.zeros(N, 32, device=dev)
        c = torch.zeros(N, 32, device=dev)
        ae = emb0
        results, logits = [], []
        result = None
        for i in range(MAX_SELECTED_UNITS_NUM):
            if i > 0:
                if i == 1:
                    step_mask[rows, entity_num] = True
                step_mask[rows, result] = False
            q = self.fc(pre + 'query_fc2', self.fc(pre + 'query_fc1', ae, relu=True))
            h, c = self.lstm_cell(cp, self.ln(cp + '.layernorm_i', q @ self.P[cp + '.weight_ih'].t()), h, c)
            step_logits = (h.unsqueeze(1) * key).sum(dim=2).masked_fill(~step_mask, -1e9) / self.T
            result = self.sample(step_logits)
            num = torch.where((result == entity_num) & ~end_flag, torch.full_like(num, i + 1), num)
            end_flag = end_flag | (result == entity_num)
            results.append(result)
            logits.append(step_logits)
            picked[rows[~end_flag], result[~end_flag]] = 1
            ae = emb0 + self.su_embed(key, picked, (picked.sum(dim=1, keepdim=True) != 0))
            if bool(end_flag.all()):
                break
        extra = torch.zeros(N, MAX_ENTITY_NUM + 1, device=dev)
        return torch.stack(logits, dim=1), torch.stack(results, dim=1), ae, num, extra

    def target_unit_head(self, emb, entity_embeddings, entity_num, target_unit=None):
        """action_arg_head.py:343-363 (K13)."""
        pre = 'policy.target_unit_head.'
        key = self.head_keys(entity_embeddings)[1]
        q = self.fc(pre + 'query_fc2', self.fc(pre + 'query_fc1', emb, relu=True))
        if self._head_keys_full is not None:      # K13: dot + mask + temperature as one warp-level kernel on the stacked keys
            logits = ops.target_unit_logits(self._head_keys_full, 32, q, entity_num, self.T, self._head_keys_sink)
        else:
            logits = ops.target_unit_logits(key.contiguous(), 0, q, entity_num, self.T)
        if target_unit is None:
            target_unit = self.sample(logits)
        return logits, target_unit

    def location_head(self, emb, map_skip, location=None):
        """action_arg_head.py:417-450 (K14).  map_skip[3..6] are channels-last [P,16,16,128]."""
        pre = 'policy.location_head.'
        P_ = self.P
        N = emb.shape[0]
        h8, w8 = map_skip[-1].shape[1:3]
        x4 = self.fc(pre + 'project_embed', emb, relu=True).reshape(N, 4, h8, w8).permute(0, 2, 3, 1)   # [N,16,16,4]
        # conv1 (1x1 over cat[x4, skip]) = W[:, :4] . relu(x4)  +  W[:, 4:] . relu(skip): the 4-channel part is a tiny
        # matmul handed to the 128-channel implicit GEMM as its residual
        w1 = P_[pre + 'conv1.0.weight']
        small = torch.matmul(torch.relu(x4), w1[:, :4, 0, 0].t())
        x = ops.conv_nhwc(torch.relu(map_skip[-1]), w1[:, 4:], P_[pre + 'conv1.0.bias'], True, small, self.terms)
        x = x + map_skip[len(map_skip) - 1]
        if x.is_cuda:
            x = ops.attach_split(x, *ops.split_bf16(x))            # conv1 and the first gate conv both read it
        for i in range(4):
            rp = pre + 'res.%d.' % i                                   # GatedResBlock, module_utils.py:224-231
            r = self.conv_nhwc(rp + 'conv2', self.conv_nhwc(rp + 'conv1', x, relu=True, split='only'))
            g = x
            for j in range(4):
                g = self.conv_nhwc(rp + 'GateWeightG.%d' % j, g, relu=(j < 3), split=('only' if j < 3 else False))
            # gate + residual + ReLU, and the `x + map_skip` that opens the next block, in one pass
            x = ops.gate_update(r, g, x, P_[rp + 'UpdateSP'], map_skip[len(map_skip) - i - 2] if i < 3 else None)
        # up-sampling stages: the 3x3 taps are applied as nine shifted up-samplings of a low-resolution projection
        x = ops.upsample_conv3x3(x, P_[pre + 'upsample.0.0.weight'], P_[pre + 'upsample.0.0.bias'], True,
                                 pair_only=True)                                                       # [N,32,32,64]
        x = ops.upsample_conv3x3(x, P_[pre + 'upsample.1.0.weight'], P_[pre + 'upsample.1.0.bias'], True)   # [N,64,64,32]
        # last stage (32 -> 1 channel at 128x128): up-sampling commutes with the channel contraction
        x = ops.upsample_conv3x3_single(x, P_[pre + 'upsample.2.0.weight'], P_[pre + 'upsample.2.0.bias'])
        logits = x.reshape(N, -1) / self.T
        if location is None:
            location = self.sample(logits)
        return logits, location

    def value_baseline(self, name: str, x: Tensor) -> Tensor:
        """model/value.py:31-39 (K16)."""
        pre = 'value_networks.%s.' % name
        x = self.fc(pre + 'project', x, relu=True)
        for i in range(16):
            r = self.fc(pre + 'res.%d.fc2' % i, self.fc(pre + 'res.%d.fc1' % i, x, relu=True, split='only'))
            x = self.ln(pre + 'res.%d.norm' % i, r, residual=x, split=True)
        v = self.fc(pre + 'value_fc', x).squeeze(1)
        if BASELINE_ATAN[name]:
            v = (2.0 / math.pi) * torch.atan((math.pi / 2.0) * v)
        return v

    # -------------------------------------------------------------------------------------- value encoder
    def value_encoder_flat(self, vf: Dict[str, Tensor]):
        """ValueEncoder.forward, obs_encoder/value_encoder.py:47-63,73: everything but the spatial tower, over all rows at once.
        Returns (fc embeddings [N,352], beginning-order embedding [N,64], masked scatter projection [N,512,8])."""
        P, pre = self.P, 'value_encoder.'
        em = pre + 'encode_modules.'
        fc = []
        for k in VALUE_FC_FIELDS:                  # config order: actor_critic_default_config.yaml:29-64
            x = vf[k]
            fc.append(self.fc(em + k, x if x.is_cuda else x.float(), relu=True,
                              exact=x.dtype in (torch.uint8, torch.int8, torch.int16)))
        unit = torch.cat([F.embedding(vf['unit_alliance'].long(), P[em + 'unit_alliance.weight']),
                          F.embedding(vf['unit_type'].long(), P[em + 'unit_type.weight'])], dim=-1)    # trainable nn.Embedding
        bo = self.bo_embedding(em + 'beginning_order.', vf['beginning_order'], vf['bo_location'])
        project = self.fc(pre + 'scatter_project', unit, relu=True)
        E = project.shape[1]
        valid = torch.arange(E, device=project.device).unsqueeze(0) < vf['total_unit_count'].unsqueeze(1)
        return torch.cat(fc, dim=-1), bo, project * valid.unsqueeze(-1)

    def value_encoder_spatial(self, project: Tensor, unit_x: Tensor, unit_y: Tensor, unit_count: Tensor, own: Tensor,
                              enemy: Tensor) -> Tensor:
        """value_encoder.py:64-72: scatter_connection of the 8-wide unit projection, + the two unit-presence planes, 1x1
        project, (max-pool, 3x3 conv) x 3, four ResBlocks, spatial_fc -> [N,128].  The scatter runs on the 32-channel
        scatter_connection kernel (channels 8..31 zero); the 16/32-channel convolutions are fp32 library calls (this tower
        is 0.08 GFLOP per observation, outside the BASELINE hot path)."""
        P, pre = self.P, 'value_encoder.'
        N, _c, H, W = own.shape
        smap = ops.scatter_connection(F.pad(project, (0, 32 - project.shape[-1])), unit_x, unit_y, unit_count, H, W)
        x = torch.cat([smap[:, :project.shape[-1]], own.float(), enemy.float()], dim=1)
        x = self.conv(pre + 'project', x, 0, relu=True)
        for i in range(3):
            x = self.conv(pre + 'downsample.%d' % (2 * i + 1), F.max_pool2d(x, 2, 2), 1, relu=True)
        for i in range(4):
            r = self.conv(pre + 'res.%d.conv1' % i, x, 1, relu=True)
            x = torch.relu(self.conv(pre + 'res.%d.conv2' % i, r, 1) + x)
        return self.fc(pre + 'spatial_fc', x.reshape(N, -1), relu=True)

    def value_encoder(self, vf: Dict[str, Tensor], spatial_fn=None) -> Tensor:
        """-> [N,544] = [fc embeddings | spatial | beginning order] (value_encoder.py:73)."""
        fc, bo, project = self.value_encoder_flat(vf)
        run = spatial_fn or self.value_encoder_spatial
        sp = run(project, vf['unit_x'], vf['unit_y'], vf['total_unit_count'], vf['own_units_spatial'], vf['enemy_units_spatial'])
        return torch.cat([fc, sp, bo], dim=-1)

    # -------------------------------------------------------------------------------------- policy
    def policy_sample(self, lstm_out, entity_embeddings, map_skip, scalar_context, entity_num, su_action_mask,
                      su_fixed_steps: Optional[int] = None):
        """model/policy.py:22-48."""
        logit, action = {}, {}
        logit['action_type'], action['action_type'], emb = self.action_type_head(lstm_out, scalar_context)
        logit['delay'], action['delay'], emb = self.arg_head('policy.delay_head.', emb, 128, False)
        logit['queued'], action['queued'], emb = self.arg_head('policy.queued_head.', emb, 2, True)
        su_mask = su_action_mask.to(emb.device)[action['action_type']]
        logit['selected_units'], action['selected_units'], emb, su_num, extra = self.selected_units_sample(
            emb, entity_embeddings, entity_num, su_mask, su_fixed_steps)
        logit['target_unit'], action['target_unit'] = self.target_unit_head(emb, entity_embeddings, entity_num)
        logit['target_location'], action['target_location'] = self.location_head(emb, map_skip)
        return action, su_num, logit, extra

    def policy_train(self, lstm_out, entity_embeddings, map_skip, scalar_context, entity_num, action_info,
                     selected_units_num, su_steps: Optional[int] = None):
        """model/policy.py:50-73."""
        logit, action = {}, {}
        logit['action_type'], action['action_type'], emb = self.action_type_head(
            lstm_out, scalar_context, action_info['action_type'])
        logit['delay'], action['delay'], emb = self.arg_head('policy.delay_head.', emb, 128, False,
                                                             action_info['delay'])
        logit['queued'], action['queued'], emb = self.arg_head('policy.queued_head.', emb, 2, True,
                                                               action_info['queued'])
        logit['selected_units'], emb, su_num = self.selected_units_train(
            emb, entity_embeddings, entity_num, selected_units_num, action_info['selected_units'], su_steps)
        action['selected_units'] = None   # the reference returns None here (action_arg_head.py:166,314)
        logit['target_unit'], action['target_unit'] = self.target_unit_head(
            emb, entity_embeddings, entity_num, action_info['target_unit'])
        logit['target_location'], action['target_location'] = self.location_head(
            emb, map_skip, action_info['target_location'])
        return action, su_num, logit
