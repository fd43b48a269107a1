"""Learner-batch assembly on the GPU from compact trajectories.

Reference: ``collate_fn`` + ``padding_entity_info`` (rl_training/rl_dataloader.py:45-76,206-245) pad every trajectory step on
the host — entity fields to the batch maximum, selected-units / target-unit teacher logits and behaviour log-probs with -1e9
to [64, 513] / [512] / [64], three sequence masks — and ``to_device`` then ships the PADDED batch: 1.47 GB per step at
B=128 x T=32, of which 0.54 GB is the [T,B,64,513] selected-units teacher tensor that is -1e9 almost everywhere.

Here the wire format keeps what the actor actually produced:
  * entity fields un-padded: one [fields, total_entities] block per dtype + ``entity_num``;
  * selected-units teacher logits as [sum_r num_r * (entity_num_r + 1)] floats, target-unit teacher logits as
    [sum_r entity_num_r], selected-units behaviour log-probs / labels as [sum_r num_r];
  * no masks (they are functions of ``selected_units_num`` / ``entity_num``);
  * the six categorical spatial planes bit-packed into one uint16 per pixel next to the uint8 height map;
  * only the first B rows of the LSTM state (the learner reads nothing else, model.py:117-119);
and ``expand_rl_batch`` rebuilds the reference layout in HBM with three kernels (csrc/batch_expand.cu).  The result is
bit-identical to the padded batch (tests/test_batch_assembly.py), with entities padded to MAX_ENTITY_NUM = 512 instead of the
batch maximum (padded entities are masked everywhere, so outputs do not change).

``compact_rl_batch`` is the inverse, the packing an actor-side sender performs; here it serves the tests and bench.py.
"""
from typing import Dict

import torch

from . import lib, ops
from .policy_net import MAX_ENTITY_NUM, MAX_SELECTED_UNITS_NUM
from .synth import tree_map

PACKED_PLANES = [('visibility_map', 0, 2), ('creep', 2, 1), ('player_relative', 3, 3), ('alerts', 6, 1), ('pathable', 7, 1),
                 ('buildable', 8, 1)]                      # (name, first bit, bits)
NEG = -1e9


def compact_rl_batch(batch: Dict) -> Dict:
    """Padded learner batch (rl_dataloader.py layout, CPU tensors) -> compact wire format."""
    B, T = batch['batch_size'], batch['unroll_len']
    en_all = batch['entity_num'].long()
    en = en_all[:T * B]
    num = batch['selected_units_num'].reshape(-1).long()
    sp = batch['spatial_info']
    packed = torch.zeros_like(sp['visibility_map'], dtype=torch.int32)
    for name, bit, _bits in PACKED_PLANES:
        packed |= sp[name].to(torch.int32) << bit
    out = {'batch_size': B, 'unroll_len': T, 'entity_num': en_all.clone(), 'selected_units_num': batch['selected_units_num'].clone(),
           'height_map': sp['height_map'].clone(), 'planes_packed': packed.to(torch.int16),
           'effects': {k: v.clone() for k, v in sp.items() if k.startswith('effect_')},
           'scalar_info': {k: v.clone() for k, v in batch['scalar_info'].items()},
           'hidden_state': [(h.reshape(-1, B, h.shape[-1])[0].clone(), c.reshape(-1, B, c.shape[-1])[0].clone())
                            for h, c in batch['hidden_state']],
           'reward': {k: v.clone() for k, v in batch['reward'].items()}, 'step': batch['step'].clone(),
           'actions_mask': {k: v.clone() for k, v in batch['mask']['actions_mask'].items()},
           'frame_masks': {k: batch['mask'][k].clone() for k in ('cum_action_mask', 'build_order_mask', 'built_unit_mask', 'effect_mask')}}
    if 'value_feature' in batch:                 # use_value_feature: already wire-sized (uint8 / bool / int16 fields), passed through
        out['value_feature'] = {k: v.clone() for k, v in batch['value_feature'].items()}
    # ---- entity fields, un-padded, grouped by dtype
    keep = torch.arange(MAX_ENTITY_NUM).unsqueeze(0) < en_all.unsqueeze(1)              # [N, 512]
    groups: Dict[torch.dtype, list] = {}
    for k, v in batch['entity_info'].items():
        groups.setdefault(v.dtype, []).append(k)
    out['entity_fields'] = {str(dt): {'names': names, 'data': torch.stack([batch['entity_info'][k][keep] for k in names])}
                            for dt, names in groups.items()}
    # ---- actions / behaviour log-probs / teacher logits
    act, mu, tl = batch['action_info'], batch['behaviour_logp'], batch['teacher_logit']
    out['action_info'] = {k: v.clone() for k, v in act.items() if k != 'selected_units'}
    out['behaviour_logp'] = {k: v.clone() for k, v in mu.items() if k != 'selected_units'}
    out['teacher_logit'] = {k: v.clone() for k, v in tl.items() if k not in ('selected_units', 'target_unit')}
    step_valid = torch.arange(MAX_SELECTED_UNITS_NUM).unsqueeze(0) < num.unsqueeze(1)    # [TB, 64]
    out['su_labels'] = act['selected_units'].reshape(-1, MAX_SELECTED_UNITS_NUM)[step_valid].to(torch.int16)
    out['su_behaviour_logp'] = mu['selected_units'].reshape(-1, MAX_SELECTED_UNITS_NUM)[step_valid].clone()
    ent_valid = torch.arange(MAX_ENTITY_NUM).unsqueeze(0) < en.unsqueeze(1)              # [TB, 512]
    out['tu_teacher'] = tl['target_unit'].reshape(-1, MAX_ENTITY_NUM)[ent_valid].clone()
    slot_valid = torch.arange(MAX_ENTITY_NUM + 1).unsqueeze(0) < (en + 1).unsqueeze(1)  # [TB, 513]
    su_t = tl['selected_units'].reshape(-1, MAX_SELECTED_UNITS_NUM, MAX_ENTITY_NUM + 1)
    out['su_teacher'] = su_t[step_valid.unsqueeze(-1) & slot_valid.unsqueeze(1)].clone()
    return out


def nbytes(tree) -> int:
    n = [0]
    tree_map(lambda t: n.__setitem__(0, n[0] + t.numel() * t.element_size()) or t, tree)
    return n[0]


def _offsets(counts: torch.Tensor) -> torch.Tensor:
    return torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])


def _expand(src: torch.Tensor, off: torch.Tensor, steps, width: torch.Tensor, rows: int, S: int, W: int, fill, dtype=None):
    """dst[r, s, e] = src[off[r] + s * width[r] + e] inside (steps[r], width[r]), `fill` outside."""
    dst = torch.empty((rows, S, W), dtype=src.dtype, device=src.device)
    if src.numel() == 0:
        src = src.new_zeros(1)                 # nothing to read (e.g. no unit selections in the whole batch): all fill
    if ops._use_kernel(src):
        lib.call('dsb_expand_ragged', src, off, steps, width, dst, rows, S, W, src.element_size(), float(fill),
                 1 if src.dtype == torch.float32 else 0)
        return dst
    return ops._standin('expand_ragged')(src, off, steps, width, rows, S, W, fill)


def _seq_mask(lengths: torch.Tensor, add: int, W: int) -> torch.Tensor:
    if ops._use_kernel(lengths):
        dst = torch.empty((lengths.numel(), W), dtype=torch.uint8, device=lengths.device)
        lib.call('dsb_sequence_mask', lengths.contiguous(), add, dst, lengths.numel(), W)
        return dst.bool()
    return ops._standin('sequence_mask')(lengths, add, W)


def expand_rl_batch(compact: Dict, device=None, staged: Dict = None) -> Dict:
    """Compact wire format -> the padded batch ``Model.rl_learner_forward`` consumes, assembled on `device`.
    ``staged``: the compact payload already on the device (bench.py copies it on a side stream); otherwise it is moved now."""
    c = staged if staged is not None else tree_map(lambda t: t.to(device, non_blocking=True), compact)
    B, T = compact['batch_size'], compact['unroll_len']
    en_all = c['entity_num']
    N = en_all.numel()
    en = en_all[:T * B]
    num = c['selected_units_num'].reshape(-1).long()
    i32 = lambda t: t.to(torch.int32).contiguous()
    # ---- spatial planes
    sp = {'height_map': c['height_map']}
    hw = c['planes_packed']
    if ops._use_kernel(hw):
        planes = [torch.empty(hw.shape, dtype=torch.uint8, device=hw.device) for _ in PACKED_PLANES]
        lib.call('dsb_unpack_planes', hw.contiguous(), *planes, hw.numel())
        for (name, _, _), p in zip(PACKED_PLANES, planes):
            sp[name] = p
    else:
        sp.update(ops._standin('unpack_planes')(hw, PACKED_PLANES))
    sp.update(c['effects'])
    # ---- entity fields: zero-padded to MAX_ENTITY_NUM
    ent_off = _offsets(en_all)
    total = int(compact['entity_num'].sum())
    entity_info = {}
    for _dt, grp in c['entity_fields'].items():
        names, data = grp['names'], grp['data']
        F = len(names)
        off = (ent_off.unsqueeze(0) + total * torch.arange(F, device=ent_off.device).unsqueeze(1)).reshape(-1).contiguous()
        width = i32(en_all.repeat(F))
        full = _expand(data.reshape(-1), off, None, width, F * N, 1, MAX_ENTITY_NUM, 0).view(F, N, MAX_ENTITY_NUM)
        for i, k in enumerate(names):
            entity_info[k] = full[i]
    # ---- ragged action / teacher tensors (rl_dataloader.py:215-232)
    TB = T * B
    su_off = _offsets(num)
    labels = _expand(c['su_labels'], su_off, None, i32(num), TB, 1, MAX_SELECTED_UNITS_NUM, 0).view(T, B, -1).long()
    mu_su = _expand(c['su_behaviour_logp'], su_off, None, i32(num), TB, 1, MAX_SELECTED_UNITS_NUM, NEG).view(T, B, -1)
    tu = _expand(c['tu_teacher'], _offsets(en), None, i32(en), TB, 1, MAX_ENTITY_NUM, NEG).view(T, B, -1)
    w1 = en + 1
    su_t = _expand(c['su_teacher'], _offsets(num * w1), i32(num), i32(w1), TB, MAX_SELECTED_UNITS_NUM, MAX_ENTITY_NUM + 1,
                   NEG).view(T, B, MAX_SELECTED_UNITS_NUM, MAX_ENTITY_NUM + 1)
    action_info = dict(c['action_info'], selected_units=labels)
    behaviour_logp = dict(c['behaviour_logp'], selected_units=mu_su)
    teacher = dict(c['teacher_logit'], selected_units=su_t, target_unit=tu)
    mask = dict(c['frame_masks'])
    mask['actions_mask'] = c['actions_mask']
    mask['selected_units_mask'] = _seq_mask(num, 0, MAX_SELECTED_UNITS_NUM).view(T, B, -1)
    mask['selected_units_logits_mask'] = _seq_mask(en, 1, MAX_ENTITY_NUM + 1).view(T, B, -1)
    mask['target_units_logits_mask'] = _seq_mask(en, 0, MAX_ENTITY_NUM).view(T, B, -1)
    out = {'spatial_info': sp, 'entity_info': entity_info, 'scalar_info': c['scalar_info'], 'entity_num': en_all,
           'hidden_state': c['hidden_state'], 'action_info': action_info, 'selected_units_num': c['selected_units_num'],
           'behaviour_logp': behaviour_logp, 'teacher_logit': teacher, 'mask': mask, 'reward': c['reward'], 'step': c['step'],
           'batch_size': B, 'unroll_len': T}
    if 'value_feature' in c:
        out['value_feature'] = c['value_feature']
    return out
