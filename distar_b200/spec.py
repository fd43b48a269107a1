"""Parameter inventory of the AlphaStar policy as plain data.

The names and shapes are the reference's ``Model.state_dict()`` keys (DI-star
``distar/agent/default/model/model.py:23-44`` and submodules) so checkpoints interchange with the
reference: actors strip ``value_networks*`` keys (``distar/actor/actor.py:71-73``) and the learner
filters them when publishing weights (``ctools/worker/learner/learner_comm.py:74``), so the key names are
API surface.  ``kind`` tells ``params.init_state_dict`` how to draw the tensor and tells the flat arena
whether the entry is trainable ("frozen" = the non-trainable one-hot / binary lookup tables the reference
registers as ``nn.Embedding.from_pretrained(..., freeze=True)``; they are kept only for checkpoint
compatibility — no kernel ever reads them).
"""
from typing import List, Tuple

Spec = Tuple[str, Tuple[int, ...], str]

BASELINES = ['winloss', 'build_order', 'built_unit', 'effect', 'upgrade', 'battle']

# entity one-hot / binary tables: (name, rows, cols)   entity_encoder.py:35-39
ENTITY_TABLES = [
    ('unit_type', 260, 260), ('alliance', 5, 5), ('cargo_space_taken', 9, 9), ('display_type', 5, 5),
    ('x', 2048, 11), ('y', 2048, 11), ('cloak', 5, 5), ('is_blip', 2, 2), ('is_powered', 2, 2),
    ('cargo_space_max', 9, 9), ('assigned_harvesters', 24, 24), ('weapon_cooldown', 32, 32),
    ('order_length', 9, 9), ('order_id_0', 327, 327), ('order_id_1', 49, 49), ('is_hallucination', 2, 2),
    ('buff_id_0', 50, 50), ('buff_id_1', 50, 50), ('addon_unit_type', 9, 9), ('is_active', 2, 2),
    ('order_id_2', 49, 49), ('order_id_3', 49, 49), ('is_in_cargo', 2, 2), ('attack_upgrade_level', 4, 4),
    ('armor_upgrade_level', 4, 4), ('shield_upgrade_level', 4, 4), ('last_selected_units', 2, 2),
    ('last_targeted_unit', 2, 2),
]


def _fc(out: List[Spec], name: str, din: int, dout: int):
    out.append((name + '.0.weight', (dout, din), 'xavier_normal'))
    out.append((name + '.0.bias', (dout,), 'bias:%d' % din))


def _conv(out: List[Spec], name: str, cin: int, cout: int, k: int):
    out.append((name + '.0.weight', (cout, cin, k, k), 'xavier_normal'))
    out.append((name + '.0.bias', (cout,), 'bias:%d' % (cin * k * k)))


def _ln(out: List[Spec], name: str, d: int):
    out.append((name + '.weight', (d,), 'ln_weight'))
    out.append((name + '.bias', (d,), 'ln_bias'))


def _transformer(out: List[Spec], pre: str, din: int, dim: int, hidden: int, heads: int, head_dim: int):
    _fc(out, pre + '.embedding', din, dim)
    for i in range(3):
        lp = '%s.layers.%d' % (pre, i)
        _fc(out, lp + '.attention.attention_pre', dim, heads * head_dim * 3)
        _fc(out, lp + '.attention.project', heads * head_dim, dim)
        _ln(out, lp + '.layernorm1', dim)
        _fc(out, lp + '.mlp.0', dim, hidden)
        _fc(out, lp + '.mlp.1', hidden, dim)
        _ln(out, lp + '.layernorm2', dim)


def _lnlstm(out: List[Spec], pre: str, din: int, hid: int, layers: int):
    for l in range(layers):
        cp = '%s.layers.%d.cell' % (pre, l)
        out.append((cp + '.weight_ih', (4 * hid, din if l == 0 else hid), 'randn'))
        out.append((cp + '.weight_hh', (4 * hid, hid), 'randn'))
        _ln(out, cp + '.layernorm_i', 4 * hid)
        _ln(out, cp + '.layernorm_h', 4 * hid)
        _ln(out, cp + '.layernorm_c', hid)


VALUE_FEATURE_DIM = 544          # ValueEncoder output: 352 (fc embeddings) + 128 (spatial_fc) + 64 (beginning order)


def value_encoder_specs(o: List[Spec], spatial_x: int, spatial_y: int) -> None:
    """ValueEncoder (obs_encoder/value_encoder.py:12-45; dims actor_critic_default_config.yaml:27-74), registration order."""
    ve = 'value_encoder.'
    em = ve + 'encode_modules.'
    _fc(o, em + 'enemy_unit_counts_bow', 260, 64)
    _fc(o, em + 'enemy_unit_type_bool', 260, 64)
    _fc(o, em + 'enemy_agent_statistics', 10, 64)
    _fc(o, em + 'enemy_upgrades', 90, 32)
    o.append((em + 'unit_alliance.weight', (2, 16), 'randn'))            # nn.Embedding, trainable, N(0, 1)
    o.append((em + 'unit_type.weight', (260, 48), 'randn'))
    _fc(o, em + 'cumulative_stat', 167, 128)
    _transformer(o, em + 'beginning_order.transformer', 214, 64, 128, 2, 8)
    _fc(o, em + 'beginning_order.embedd_fc', 64, 64)
    o.append((em + 'beginning_order.action_one_hot.weight', (174, 174), 'frozen_eye'))
    o.append((em + 'beginning_order.order_one_hot.weight', (20, 20), 'frozen_eye'))
    o.append((em + 'beginning_order.location_binary.weight', (1024, 10), 'frozen_binary'))
    _fc(o, ve + 'scatter_project', 64, 8)
    _conv(o, ve + 'project', 10, 16, 1)
    for i, (a, b) in enumerate([(16, 16), (16, 32), (32, 32)]):
        _conv(o, ve + 'downsample.%d' % (2 * i + 1), a, b, 3)            # nn.Sequential(pool, conv, pool, conv, pool, conv)
    for i in range(4):
        _conv(o, ve + 'res.%d.conv1' % i, 32, 32, 3)
        _conv(o, ve + 'res.%d.conv2' % i, 32, 32, 3)
    _fc(o, ve + 'spatial_fc', 32 * (spatial_y // 8) * (spatial_x // 8), 128)


def param_specs(spatial_x: int = 128, spatial_y: int = 128, baselines=('winloss',), use_value_feature: bool = False) -> List[Spec]:
    """Ordered exactly like the reference's state_dict."""
    o: List[Spec] = []
    se = 'encoder.scalar_encoder.'
    o.append((se + 'position_array', (32,), 'position_array'))
    em = se + 'encode_modules.'
    _fc(o, em + 'agent_statistics', 10, 64)
    o.append((em + 'home_race.weight', (5, 32), 'xavier_uniform'))
    o.append((em + 'away_race.weight', (5, 32), 'xavier_uniform'))
    _fc(o, em + 'upgrades', 90, 128)
    _fc(o, em + 'unit_counts_bow', 260, 128)
    o.append((em + 'last_delay.weight', (128, 64), 'xavier_uniform'))
    o.append((em + 'last_queued.weight', (2, 32), 'xavier_uniform'))
    o.append((em + 'last_action_type.weight', (327, 128), 'xavier_uniform'))
    _fc(o, em + 'cumulative_stat', 167, 128)
    _fc(o, em + 'unit_type_bool', 260, 64)
    _fc(o, em + 'enemy_unit_type_bool', 260, 64)
    _fc(o, em + 'unit_order_type', 269, 64)
    _transformer(o, em + 'beginning_order.transformer', 214, 64, 128, 2, 8)
    _fc(o, em + 'beginning_order.embedd_fc', 64, 64)
    o.append((em + 'beginning_order.action_one_hot.weight', (174, 174), 'frozen_eye'))
    o.append((em + 'beginning_order.order_one_hot.weight', (20, 20), 'frozen_eye'))
    o.append((em + 'beginning_order.location_binary.weight', (1024, 10), 'frozen_binary'))
    sp = 'encoder.spatial_encoder.'
    _conv(o, sp + 'project', 56, 32, 1)
    for n, k in [('visibility_map', 4), ('creep', 2), ('player_relative', 5), ('alerts', 2), ('pathable', 2),
                 ('buildable', 2)]:
        o.append((sp + 'encode_modules.%s.weight' % n, (k, k), 'frozen_eye'))
    for i, (a, b) in enumerate([(32, 64), (64, 128), (128, 128)]):
        _conv(o, sp + 'downsample.%d' % i, a, b, 3)
    for i in range(4):
        _conv(o, sp + 'res.%d.conv1' % i, 128, 128, 3)
        _conv(o, sp + 'res.%d.conv2' % i, 128, 128, 3)
    _fc(o, sp + 'fc', 128 * (spatial_y // 8) * (spatial_x // 8), 256)
    ee = 'encoder.entity_encoder.'
    for n, r, c in ENTITY_TABLES:
        o.append((ee + 'encode_modules.%s.weight' % n, (r, c), 'frozen_binary' if n in ('x', 'y') else 'frozen_eye'))
    _transformer(o, ee + 'transformer', 997, 256, 1024, 2, 128)
    _fc(o, ee + 'entity_fc', 256, 256)
    _fc(o, ee + 'embed_fc', 256, 256)
    _fc(o, 'encoder.scatter_project', 256, 32)
    at = 'policy.action_type_head.'
    _fc(o, at + 'project', 384, 256)
    for i in range(2):
        for j in (1, 2):
            _fc(o, at + 'res.%d.fc%d' % (i, j), 256, 256)
            o.append((at + 'res.%d.fc%d.1.weight' % (i, j), (256,), 'ln_weight'))
            o.append((at + 'res.%d.fc%d.1.bias' % (i, j), (256,), 'ln_bias'))
    for n, ctx, din, dout in [('action_fc', 448, 256, 327)]:
        _fc(o, at + n + '.layer1', ctx, din)
        _fc(o, at + n + '.layer2', din, dout)
    _fc(o, at + 'action_map_fc1', 327, 256)
    _fc(o, at + 'action_map_fc2', 256, 256)
    for n, din in [('glu1', 256), ('glu2', 384)]:
        _fc(o, at + n + '.layer1', 448, din)
        _fc(o, at + n + '.layer2', din, 1024)
    for hn, n in [('delay_head', 128), ('queued_head', 2)]:
        hp = 'policy.%s.' % hn
        _fc(o, hp + 'fc1', 1024, 256)
        _fc(o, hp + 'fc2', 256, 256)
        _fc(o, hp + 'fc3', 256, n)
        _fc(o, hp + 'embed_fc1', n, 256)
        _fc(o, hp + 'embed_fc2', 256, 1024)
    su = 'policy.selected_units_head.'
    o.append((su + 'end_embedding', (1, 32), 'uniform:32'))
    _fc(o, su + 'key_fc', 256, 32)
    _fc(o, su + 'query_fc1', 1024, 256)
    _fc(o, su + 'query_fc2', 256, 32)
    _fc(o, su + 'embed_fc1', 32, 256)
    _fc(o, su + 'embed_fc2', 256, 1024)
    _lnlstm(o, su + 'lstm', 32, 32, 1)
    tu = 'policy.target_unit_head.'
    _fc(o, tu + 'key_fc', 256, 32)
    _fc(o, tu + 'query_fc1', 1024, 32)
    _fc(o, tu + 'query_fc2', 32, 32)
    lh = 'policy.location_head.'
    _conv(o, lh + 'conv1', 132, 128, 1)
    for i in range(4):
        o.append((lh + 'res.%d.UpdateSP' % i, (1,), 'const:0.1'))
        _conv(o, lh + 'res.%d.conv1' % i, 128, 128, 3)
        _conv(o, lh + 'res.%d.conv2' % i, 128, 128, 3)
        for j in range(4):
            _conv(o, lh + 'res.%d.GateWeightG.%d' % (i, j), 128, 128, 1)
    _fc(o, lh + 'project_embed', 1024, (spatial_y // 8) * (spatial_x // 8) * 4)
    for i, (a, b) in enumerate([(128, 64), (64, 32), (32, 1)]):
        _conv(o, lh + 'upsample.%d' % i, a, b, 3)
    if use_value_feature and len(baselines):          # model.py:31-34: only with a value network
        value_encoder_specs(o, spatial_x, spatial_y)
    for b in BASELINES:
        if b not in baselines:
            continue
        vp = 'value_networks.%s.' % b
        # value.py:20-23: + 1056 = value feature (544) + the scalar encoder's baseline feature (512)
        _fc(o, vp + 'project', 384 + (VALUE_FEATURE_DIM + 512 if use_value_feature else 0), 256)
        for i in range(16):
            _fc(o, vp + 'res.%d.fc1' % i, 256, 256)
            _fc(o, vp + 'res.%d.fc2' % i, 256, 256)
            _ln(o, vp + 'res.%d.norm' % i, 256)
        o.append((vp + 'value_fc.0.weight', (1, 256), 'xavier_uniform:0.1'))
        o.append((vp + 'value_fc.0.bias', (1,), 'zeros'))
    _lnlstm(o, 'core_lstm', 1536, 384, 3)
    return o


def is_trainable(kind: str) -> bool:
    return not (kind.startswith('frozen') or kind == 'position_array')
