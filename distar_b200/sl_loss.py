"""``SupervisedLoss`` — drop-in for DI-star ``sl_training/sl_loss.py:37-286`` (six masked cross-entropies + metrics).

``SupervisedLoss(cfg).compute_loss(policy_logits, actions, actions_mask, selected_units_num, entity_num,
infer_action_info)`` returns the same keys ('<head>_loss', 'total_loss' and the no-grad metrics registered in
sl_loss.py:69-98).  Each head's cross-entropy is the fused per-row kernel (ops.categorical_stats): the [N*S, 513] /
[N, 16384] logits are read once, no softmax tensor is materialised; the label-smoothing term (mean_j log p_j,
sl_loss.py:16-34) comes out of the same pass.

Options follow the reference, including its defaults: ``learner.su_mask`` (default TRUE, default_supervised_loss.yaml:10;
bin/sl_user_config.yaml:28 turns it off) masks every other labelled unit out of each pointer step's logits before the
cross-entropy (sl_loss.py:177-192); ``learner.label_smooth`` swaps in LabelSmoothingCrossEntropy(0.1) for the five
non-pointer heads (:54-57); ``learner.cross_rank_loss`` re-weights by the global batch (:101-104).
"""
from typing import Dict, Optional

import torch

from . import ops
from .dist import allreduce, get_world_size

HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']
DEFAULT_WEIGHTS = {'action_type': 30.0, 'delay': 9.0, 'queued': 1.0, 'selected_units': 4.0, 'target_unit': 4.0,
                   'target_location': 8.0}       # default_supervised_loss.yaml:2-9 == bin/sl_user_config.yaml:43-51
SMOOTHING = 0.1                                   # LabelSmoothingCrossEntropy default, sl_loss.py:19


def _get(cfg, path, default):
    cur = cfg
    for k in path.split('.'):
        if isinstance(cur, dict) and k in cur:
            cur = cur[k]
        else:
            return default
    return cur


def _sequence_mask(lengths: torch.Tensor, max_len: int) -> torch.Tensor:
    return torch.arange(max_len, device=lengths.device).unsqueeze(0) < lengths.unsqueeze(1)


class SupervisedLoss:
    def __init__(self, cfg: dict = None) -> None:
        cfg = cfg or {}
        self.loss_weight = dict(DEFAULT_WEIGHTS)
        self.loss_weight.update({k: v for k, v in _get(cfg, 'learner.loss_weight', {}).items() if k in DEFAULT_WEIGHTS})
        self.su_mask = bool(_get(cfg, 'learner.su_mask', True))
        self.label_smooth = bool(_get(cfg, 'learner.label_smooth', False))
        self.cross_rank_loss = bool(_get(cfg, 'learner.cross_rank_loss', False))
        self.world_size = get_world_size()
        self.total_batch_size: Optional[torch.Tensor] = None

    LOGGED = ['total_loss', 'action_type_loss', 'delay_loss', 'queued_loss', 'selected_units_loss', 'selected_units_loss_norm',
              'selected_units_end_flag_loss', 'target_unit_loss', 'target_location_loss', 'action_type_acc', 'delay_distance_L1',
              'queued_acc', 'selected_units_iou', 'target_unit_acc', 'target_location_distance_L2']

    def register_stats(self, record, tb_logger) -> None:
        """sl_loss.py:68-98: BaseLearner.register_stats (base_learner.py:235-236) asks the loss to register what compute_loss
        returns; the framework's variable record raises on any key it has not been told about (log_helper.py:356-364)."""
        for k in self.LOGGED:
            record.register_var(k)
            tb_logger.register_var(k)

    # ---- criteria (sl_loss.py:54-58): plain CE for the pointer head, optionally smoothed CE elsewhere
    def _ce(self, logits, labels, smooth: bool = False):
        if smooth:
            logp, _, _, mean_lp = ops.categorical_stats(logits, labels, want_mean=True)
            return (1.0 - SMOOTHING) * (-logp) + SMOOTHING * (-mean_lp)
        return -ops.categorical_stats(logits, labels)[0]

    def _reduce(self, ce, mask, n_rows):
        """sl_loss.py:126-136 (and the same block in every non-pointer head)."""
        if self.cross_rank_loss:
            return (n_rows / self.total_batch_size * self.world_size) * ce.mean()
        valid = mask.sum()
        return torch.where(valid > 0, ce.sum() / valid.clamp(min=1.0), ce.sum() * 0)

    @staticmethod
    def _mask_selected(logits, labels, lengths):
        """sl_loss.py:177-192: at every pointer step, the logits of every OTHER labelled unit (end token excluded) are set
        to -1e9; the step's own label stays."""
        b, s, n = logits.shape
        keep = _sequence_mask((lengths - 1).clamp(min=0), labels.shape[1])
        nl = torch.where(keep, labels, torch.full_like(labels, n))[:, :s]          # padded / end-token steps -> dummy column n
        chosen = torch.zeros((b, n + 1), dtype=torch.bool, device=logits.device).scatter_(1, nl, True)   # any step's label
        own = torch.zeros((b, s, n + 1), dtype=torch.bool, device=logits.device).scatter_(2, nl.unsqueeze(2), True)
        blocked = (chosen.unsqueeze(1) & ~own)[:, :, :n]
        return logits.masked_fill(blocked, -1e9)

    @staticmethod
    def _iou(preds, labels, select_mask, entity_num, mask, n):
        """sl_loss.py:206-232: IoU between the sampled and the labelled unit sets."""
        b, s = labels.shape
        end = (preds == entity_num.unsqueeze(1)).long().argmax(dim=-1)
        invalid = end == 0
        end = end + 1 + invalid.long() * s
        preds = (preds + 1) * _sequence_mask(end, preds.shape[1])
        labels = (labels + 1) * select_mask
        ps = torch.zeros((b, n + 1), dtype=torch.bool, device=labels.device).scatter_(1, preds.long(), True)
        ls = torch.zeros((b, n + 1), dtype=torch.bool, device=labels.device).scatter_(1, labels.long(), True)
        inter, union = (ps & ls)[:, 1:].sum(dim=1), (ps | ls)[:, 1:].sum(dim=1)
        return (inter / (union + 1e-6) * mask).sum() / (mask.sum() + 1e-6)

    def compute_loss(self, policy_logits, actions, actions_mask, selected_units_num, entity_num,
                     infer_action_info=None) -> Dict:
        if self.cross_rank_loss:                                            # sl_loss.py:101-104
            self.total_batch_size = torch.tensor(entity_num.shape[0], device=entity_num.device)
            allreduce(self.total_batch_size, reduce=False)
        out = {}
        for h in HEADS:
            mask = actions_mask[h].float()
            if h == 'selected_units':                                       # sl_loss.py:174-238
                lg = policy_logits[h]
                b, s, n = lg.shape
                labels = actions[h][:, :s]
                if self.su_mask:
                    lg = self._mask_selected(lg, actions[h], selected_units_num)
                ce = self._ce(lg, labels)
                select = _sequence_mask(selected_units_num, s)
                ce = ce.masked_fill(~select, 0) * mask.unsqueeze(1)
                loss = ce.sum() / b
                if self.cross_rank_loss:
                    loss = (b / self.total_batch_size * self.world_size) * loss
                out[h + '_loss'] = loss
                out['selected_units_loss_norm'] = (ce.sum() / (selected_units_num.sum() + 1e-6)).detach()
                rows = torch.arange(b, device=lg.device)
                out['selected_units_end_flag_loss'] = ce[rows, selected_units_num - 1].mean().detach()
                preds = infer_action_info.get('selected_units') if infer_action_info is not None else None
                with torch.no_grad():
                    out['selected_units_iou'] = self._iou(preds, labels, select, entity_num, mask, n) if preds is not None \
                        else torch.zeros((), device=lg.device)
            else:                                                           # sl_loss.py:120-172,240-286
                ce = self._ce(policy_logits[h], actions[h], self.label_smooth) * mask
                out[h + '_loss'] = self._reduce(ce, mask, actions[h].shape[0])
            with torch.no_grad():
                if h == 'action_type':
                    out['action_type_acc'] = (policy_logits[h].argmax(1) == actions[h]).float().sum() / actions[h].shape[0]
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
