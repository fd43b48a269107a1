"""Seeded synthetic observations / trajectories — the measurement fixture of SURVEY.md §8(d).

dtypes and shapes follow ``fake_step_data`` (DI-star ``distar/agent/default/lib/features.py:95-127``:
SPATIAL_INFO/SCALAR_INFO/ENTITY_INFO :41-67) at ``SPATIAL_SIZE = [128, 128]``; value ranges follow the
vocabularies in ``model/actor_critic_default_config.yaml``.  The RL batch layout is what
``rl_training/rl_dataloader.py:45-76,206-245`` (collate_fn / padding_entity_info) produces: observation keys
time-major flattened ``[(T+1)*B, ...]``, action / mask / reward keys ``[T, B, ...]``.
Everything is generated on the CPU from a ``torch.Generator`` so a batch is reproducible on any machine.
"""
from typing import Dict

import torch

from .spec import ENTITY_TABLES

H = W = 128
E = 512
S_MAX = 64
EFFECT_LEN = 100
HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']
N_CLS = {'action_type': 327, 'delay': 128, 'queued': 2, 'target_unit': E, 'target_location': H * W}

_ENTITY_DTYPES = {  # lib/features.py:57-67
    'unit_type': torch.int16, 'order_id_0': torch.int16, 'order_id_1': torch.int16, 'order_id_2': torch.int16,
    'order_id_3': torch.int16, 'last_selected_units': torch.int8, 'last_targeted_unit': torch.int8}
_ENTITY_FLOATS = ['build_progress', 'health_ratio', 'shield_ratio', 'energy_ratio', 'mineral_contents',
                  'vespene_contents', 'order_progress_0', 'order_progress_1']
_ENTITY_ORDER = ['unit_type', 'alliance', 'cargo_space_taken', 'build_progress', 'health_ratio', 'shield_ratio',
                 'energy_ratio', 'display_type', 'x', 'y', 'cloak', 'is_blip', 'is_powered', 'mineral_contents',
                 'vespene_contents', 'cargo_space_max', 'assigned_harvesters', 'weapon_cooldown', 'order_length',
                 'order_id_0', 'order_id_1', 'is_hallucination', 'buff_id_0', 'buff_id_1', 'addon_unit_type',
                 'is_active', 'order_progress_0', 'order_progress_1', 'order_id_2', 'order_id_3', 'is_in_cargo',
                 'attack_upgrade_level', 'armor_upgrade_level', 'shield_upgrade_level', 'last_selected_units',
                 'last_targeted_unit']
_VOCAB = {n: r for n, r, _ in ENTITY_TABLES}


def _ri(g, lo, hi, shape, dtype):
    return torch.randint(lo, hi, shape, generator=g).to(dtype)


def synth_obs(n: int, seed: int = 0, entity_num=None, hidden: bool = True, hw=None) -> Dict:
    """n observation rows. entity_num: None -> 512 for every row; 'random' -> U{64..512}; or a LongTensor.
    hw: (spatial_y, spatial_x) of the maps, default 128 x 128 (the reference's own default is (152, 160))."""
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    H, W = hw if hw is not None else (globals()['H'], globals()['W'])
    sp = {'height_map': _ri(g, 0, 256, (n, H, W), torch.uint8),
          'visibility_map': _ri(g, 0, 4, (n, H, W), torch.uint8),
          'creep': _ri(g, 0, 2, (n, H, W), torch.uint8),
          'player_relative': _ri(g, 0, 5, (n, H, W), torch.uint8),
          'alerts': _ri(g, 0, 2, (n, H, W), torch.uint8),
          'pathable': _ri(g, 0, 2, (n, H, W), torch.uint8),
          'buildable': _ri(g, 0, 2, (n, H, W), torch.uint8)}
    for k in ['effect_PsiStorm', 'effect_NukeDot', 'effect_LiberatorDefenderZone', 'effect_BlindingCloud',
              'effect_CorrosiveBile', 'effect_LurkerSpines']:
        cnt = torch.randint(0, 21, (n, 1), generator=g)
        idx = torch.randint(0, H * W, (n, EFFECT_LEN), generator=g)
        sp[k] = (idx * (torch.arange(EFFECT_LEN).unsqueeze(0) < cnt)).to(torch.int16)
    ent = {}
    for k in _ENTITY_ORDER:
        if k in _ENTITY_FLOATS:
            ent[k] = torch.rand((n, E), generator=g).to(torch.float16)
        elif k in ('x', 'y'):
            ent[k] = _ri(g, 0, W if k == 'x' else H, (n, E), torch.uint8)
        else:
            ent[k] = _ri(g, 0, _VOCAB[k], (n, E), _ENTITY_DTYPES.get(k, torch.uint8))
    sc = {'home_race': _ri(g, 1, 4, (n,), torch.uint8), 'away_race': _ri(g, 1, 4, (n,), torch.uint8),
          'upgrades': _ri(g, 0, 2, (n, 90), torch.int16),
          'time': torch.rand((n,), generator=g) * 30000.,
          'unit_counts_bow': _ri(g, 0, 21, (n, 260), torch.uint8),
          'agent_statistics': torch.rand((n, 10), generator=g) * 10.,
          'cumulative_stat': (torch.rand((n, 167), generator=g) < 0.1).to(torch.uint8),
          'beginning_order': _ri(g, 0, 174, (n, 20), torch.int16),
          'last_queued': _ri(g, 0, 2, (n,), torch.int16), 'last_delay': _ri(g, 0, 128, (n,), torch.int16),
          'last_action_type': _ri(g, 0, 327, (n,), torch.int16),
          'bo_location': _ri(g, 0, H * W, (n, 20), torch.int16),
          'unit_order_type': (torch.rand((n, 269), generator=g) < 0.1).to(torch.uint8),
          'unit_type_bool': (torch.rand((n, 260), generator=g) < 0.1).to(torch.uint8),
          'enemy_unit_type_bool': (torch.rand((n, 260), generator=g) < 0.1).to(torch.uint8)}
    if entity_num is None:
        en = torch.full((n,), E, dtype=torch.long)
    elif isinstance(entity_num, str):
        en = torch.randint(64, E + 1, (n,), generator=g)
    else:
        en = entity_num.long()
    out = {'spatial_info': sp, 'entity_info': ent, 'scalar_info': sc, 'entity_num': en}
    if hidden:
        out['hidden_state'] = [(torch.randn((n, 384), generator=g), torch.randn((n, 384), generator=g))
                               for _ in range(3)]
    return out


def synth_actions(rows: int, entity_num: torch.Tensor, g: torch.Generator, max_su: int = 12, hw=None):
    """Teacher-forced labels for `rows` policy rows: selected_units_num in {0} U [2, max_su] (never 1:
    the reference's teacher-forced path divides 0/0 there, action_arg_head.py:196-198)."""
    en = entity_num.long()
    a = {'action_type': torch.randint(0, 327, (rows,), generator=g),
         'delay': torch.randint(0, 128, (rows,), generator=g),
         'queued': torch.randint(0, 2, (rows,), generator=g),
         'target_unit': (torch.rand((rows,), generator=g) * en).long().clamp(max=E - 1),
         'target_location': torch.randint(0, (hw[0] * hw[1]) if hw is not None else H * W, (rows,), generator=g)}
    num = torch.randint(2, max_su + 1, (rows,), generator=g)
    num = torch.where(torch.rand((rows,), generator=g) < 0.2, torch.zeros_like(num), num)
    num = torch.minimum(num, en)          # need num-1 distinct units + end token
    num = torch.where(num == 1, torch.zeros_like(num), num)
    su = torch.zeros((rows, S_MAX), dtype=torch.long)
    keys = torch.rand((rows, E), generator=g)
    keys = keys.masked_fill(torch.arange(E).unsqueeze(0) >= en.unsqueeze(1), 2.0)
    order = keys.argsort(dim=1)[:, :S_MAX]
    pos = torch.arange(S_MAX).unsqueeze(0)
    su = torch.where(pos < (num - 1).unsqueeze(1), order, su)
    su = torch.where((pos == (num - 1).unsqueeze(1)) & (num > 0).unsqueeze(1), en.unsqueeze(1).expand(-1, S_MAX), su)
    a['selected_units'] = su
    return a, num


def _masked_logits(shape, valid: torch.Tensor, g) -> torch.Tensor:
    return torch.randn(shape, generator=g).masked_fill(~valid, -1e9)


def synth_value_feature(rows: int, seed: int = 0) -> Dict:
    """The `value_feature` entry of a learner batch (use_value_feature: True), dtypes as lib/features.py:690-765 builds them
    plus the behaviour z the agent merges in (agent.py:562-564,609-613): opponent unit statistics, the positions / types of
    all visible units (enemy first, padded to MAX_ENTITY_NUM) and two boolean unit-presence planes."""
    g = torch.Generator(device='cpu')
    g.manual_seed(seed + 15485863)
    total = _ri(g, 1, E + 1, (rows,), torch.int64)
    valid = torch.arange(E).unsqueeze(0) < total.unsqueeze(1)
    bow = _ri(g, 0, 4, (rows, 260), torch.uint8) * (torch.rand((rows, 260), generator=g) < 0.1).to(torch.uint8)
    return {'unit_type': (_ri(g, 0, 260, (rows, E), torch.int16) * valid).to(torch.int16),
            'enemy_unit_counts_bow': bow, 'enemy_unit_type_bool': (bow > 0).to(torch.uint8),
            'unit_x': (_ri(g, 0, W, (rows, E), torch.uint8) * valid).to(torch.uint8),
            'unit_y': (_ri(g, 0, H, (rows, E), torch.uint8) * valid).to(torch.uint8),
            'unit_alliance': (torch.rand((rows, E), generator=g) < 0.4) & valid, 'total_unit_count': total,
            'enemy_agent_statistics': torch.log(_ri(g, 0, 2000, (rows, 10), torch.int64).float() + 1),
            'enemy_upgrades': (torch.rand((rows, 90), generator=g) < 0.1).to(torch.uint8),
            'own_units_spatial': torch.rand((rows, 1, H, W), generator=g) < 0.02,
            'enemy_units_spatial': torch.rand((rows, 1, H, W), generator=g) < 0.02,
            'beginning_order': _ri(g, 0, 174, (rows, 20), torch.int64), 'bo_location': _ri(g, 0, H * W, (rows, 20), torch.int64),
            'cumulative_stat': (torch.rand((rows, 167), generator=g) < 0.2).long()}


def synth_rl_batch(batch_size: int, unroll_len: int, seed: int = 0, entity_num=None, max_su: int = 12,
                   value_feature: bool = False) -> Dict:
    """One learner batch in the reference's collate layout (see module docstring)."""
    B, T = batch_size, unroll_len
    obs = synth_obs((T + 1) * B, seed=seed, entity_num=entity_num)
    g = torch.Generator(device='cpu')
    g.manual_seed(seed + 7919)
    en = obs['entity_num'][:T * B]
    act, num = synth_actions(T * B, en, g, max_su)
    tb = lambda t: t.view(T, B, *t.shape[1:])
    action_info = {k: tb(v) for k, v in act.items()}
    selected_units_num = tb(num)
    su_step_valid = torch.arange(S_MAX).view(1, 1, -1) < selected_units_num.unsqueeze(-1)           # [T,B,64]
    ent_valid = torch.arange(E).view(1, 1, -1) < tb(en).unsqueeze(-1)                               # [T,B,512]
    ent1_valid = torch.arange(E + 1).view(1, 1, -1) < (tb(en) + 1).unsqueeze(-1)                    # [T,B,513]
    behaviour_logp = {k: -3.0 * torch.rand((T, B), generator=g) for k in HEADS if k != 'selected_units'}
    behaviour_logp['selected_units'] = (-3.0 * torch.rand((T, B, S_MAX), generator=g)).masked_fill(~su_step_valid, -1e9)
    teacher = {k: torch.randn((T, B, N_CLS[k]), generator=g) for k in ('action_type', 'delay', 'queued')}
    teacher['target_unit'] = _masked_logits((T, B, E), ent_valid, g)
    teacher['target_location'] = torch.randn((T, B, H * W), generator=g)
    teacher['selected_units'] = _masked_logits((T, B, S_MAX, E + 1), ent1_valid.unsqueeze(2) & su_step_valid.unsqueeze(-1), g)
    bern = lambda p, shape: (torch.rand(shape, generator=g) < p)
    mask = {'actions_mask': {k: bern(0.7, (T, B)).long() for k in ('queued', 'selected_units', 'target_unit',
                                                                   'target_location')},
            'selected_units_mask': su_step_valid.clone(),
            'selected_units_logits_mask': ent1_valid.clone(),
            'target_units_logits_mask': ent_valid.clone(),
            'cum_action_mask': bern(0.5, (T, B)).float(),
            'build_order_mask': bern(0.5, (T, B)).float(), 'built_unit_mask': bern(0.5, (T, B)).float(),
            'effect_mask': bern(0.5, (T, B)).float()}
    winloss = torch.zeros(T, B)
    term = bern(0.05, (B,))
    winloss[T - 1] = torch.where(term, torch.where(bern(0.5, (B,)), 1.0, -1.0), 0.0)
    reward = {'winloss': winloss}
    for k in ('build_order', 'built_unit', 'effect', 'upgrade', 'battle'):
        reward[k] = 0.01 * torch.randn((T, B), generator=g)
    step = torch.randint(0, 20000, (T, B), generator=g).float()
    batch = dict(obs)
    batch.update({'action_info': action_info, 'selected_units_num': selected_units_num,
                  'behaviour_logp': behaviour_logp, 'teacher_logit': teacher, 'mask': mask, 'reward': reward,
                  'step': step, 'batch_size': B, 'unroll_len': T})
    if value_feature:
        batch['value_feature'] = synth_value_feature((T + 1) * B, seed=seed)
    return batch


def tree_map(fn, x):
    if isinstance(x, torch.Tensor):
        return fn(x)
    if isinstance(x, dict):
        return {k: tree_map(fn, v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(tree_map(fn, v) for v in x)
    return x


def tree_clone(x):
    return tree_map(lambda t: t.clone(), x)


def synth_sl_batch(batch_size: int, traj_len: int, seed: int = 0, entity_num=None, max_su: int = 12) -> Dict:
    """One supervised-learning batch in the layout ``SLLearner._train`` feeds ``Model.sl_train`` (sl_learner.py:46-52,
    sl_training/sl_dataloader.py collate): observation rows BATCH-major [B*T, ...], labels per row, ``action_mask`` per head
    [B*T] float, ``traj_lens`` and the indices of trajectories that start a new episode (their LSTM state is reset)."""
    B, T = batch_size, traj_len
    obs = synth_obs(B * T, seed=seed, entity_num=entity_num, hidden=False)
    g = torch.Generator(device='cpu')
    g.manual_seed(seed + 104729)
    act, num = synth_actions(B * T, obs['entity_num'], g, max_su)
    mask = {k: (torch.ones(B * T) if k in ('action_type', 'delay') else (torch.rand(B * T, generator=g) < 0.7).float())
            for k in HEADS}
    new_episodes = [i for i in range(B) if float(torch.rand((), generator=g)) < 0.1]
    batch = dict(obs)
    batch.update({'action_info': act, 'selected_units_num': num, 'action_mask': mask, 'traj_lens': [T] * B,
                  'new_episodes': new_episodes})
    return batch
