"""``SupervisedLoss`` — drop-in for DI-star ``sl_training/sl_loss.py:37-286`` (six masked cross-entropies).

``SupervisedLoss(cfg).compute_loss(policy_logits, actions, actions_mask, selected_units_num, entity_num,
infer_action_info)`` returns the same keys ('<head>_loss', 'total_loss' and the no-grad metrics).  Each head's
cross-entropy is the fused per-row log-prob kernel (ops.categorical_stats): the [N*S, 513] / [N, 16384] logits are
read once, no softmax tensor is materialised.  ``su_mask`` (pre-masking of already selected units, sl_loss.py:177-192)
and label smoothing are off in the reference's training config (bin/sl_user_config.yaml:28, default yaml) and are not
implemented; asking for them raises.
"""
from typing import Dict

import torch

from . import ops

HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']
DEFAULT_WEIGHTS = {'action_type': 30.0, 'delay': 9.0, 'queued': 1.0, 'selected_units': 4.0, 'target_unit': 4.0,
                   'target_location': 8.0}       # bin/sl_user_config.yaml:43-51


def _get(cfg, path, default):
    cur = cfg
    for k in path.split('.'):
        if isinstance(cur, dict) and k in cur:
            cur = cur[k]
        else:
            return default
    return cur


class SupervisedLoss:
    def __init__(self, cfg: dict = None) -> None:
        cfg = cfg or {}
        self.loss_weight = dict(DEFAULT_WEIGHTS)
        self.loss_weight.update({k: v for k, v in _get(cfg, 'learner.loss_weight', {}).items() if k in DEFAULT_WEIGHTS})
        if _get(cfg, 'learner.su_mask', False):
            raise NotImplementedError('su_mask=True (sl_loss.py:177-192) is outside the benchmarked configuration')
        if _get(cfg, 'learner.label_smooth', False):
            raise NotImplementedError('label_smooth=True is outside the benchmarked configuration')

    @staticmethod
    def _ce(logits, labels):
        return -ops.categorical_stats(logits, labels)[0]

    def compute_loss(self, policy_logits, actions, actions_mask, selected_units_num, entity_num,
                     infer_action_info=None) -> Dict:
        out = {}
        for h in HEADS:
            mask = actions_mask[h].float()
            if h == 'selected_units':                                       # sl_loss.py:174-204
                lg = policy_logits[h]
                b, s, n = lg.shape
                ce = self._ce(lg, actions[h][:, :s])
                valid = torch.arange(s, device=lg.device).unsqueeze(0) < selected_units_num.unsqueeze(1)
                ce = ce.masked_fill(~valid, 0) * mask.unsqueeze(1)
                out[h + '_loss'] = ce.sum() / b
                out['selected_units_loss_norm'] = (ce.sum() / (selected_units_num.sum() + 1e-6)).detach()
                rows = torch.arange(b, device=lg.device)
                out['selected_units_end_flag_loss'] = ce[rows, selected_units_num - 1].mean().detach()
            else:                                                           # sl_loss.py:120-172,256-286
                ce = self._ce(policy_logits[h], actions[h]) * mask
                valid = mask.sum()
                out[h + '_loss'] = torch.where(valid > 0, ce.sum() / valid.clamp(min=1.0), ce.sum() * 0)
            with torch.no_grad():
                if h == 'action_type':
                    out['action_type_acc'] = (policy_logits[h].argmax(1) == actions[h]).float().mean()
                elif h == 'delay':
                    out['delay_distance_L1'] = ((policy_logits[h].argmax(-1) - actions[h]).abs() * mask).sum() / (mask.sum() + 1e-6)
                elif h == 'queued':
                    out['queued_acc'] = ((policy_logits[h].argmax(-1) - actions[h]).abs() * mask).sum() / (mask.sum() + 1e-6)
                elif h == 'target_unit':
                    out['target_unit_acc'] = ((policy_logits[h].argmax(-1) == actions[h]) * mask).sum() / (mask.sum() + 1e-6)
                elif h == 'target_location':
                    W = 160                                                 # hard-coded in the reference (sl_loss.py:257)
                    p, l = policy_logits[h].argmax(-1), actions[h]
                    d = ((p % W - l % W) ** 2 + (p // W - l // W) ** 2).float().sqrt()
                    out['target_location_distance_L2'] = (d * mask).sum() / (mask.sum() + 1e-6)
        out['total_loss'] = sum(out[h + '_loss'] * self.loss_weight[h] for h in HEADS)
        return out
