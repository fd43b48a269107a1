"""``ReinforcementLoss`` — drop-in for DI-star ``rl_training/rl_loss.py:9-199`` (V-trace PG + UPGO + TD(lambda)
critic + entropy + teacher KL + DAPO (KL towards the 'successive' model's logits, honoured only for 'MP' players and off
by default: rl_loss.py:22-24,164-172, as_rl_utils.py:105-127).

Same constructor ``ReinforcementLoss(learner_cfg, player_id)``, same ``compute_loss(model_output) -> dict`` with
'total_loss' (autograd scalar) plus the '{field}/{head}', '{field}/td', 'upgo/*', 'entropy/*', 'kl/*' floats the
learner registers (rl_learner.py:181-190), same ``only_update_value`` switch and ``reset``.

What changed underneath:
  * per head ONE fused kernel pass reads target and teacher logits once and emits log p(a), entropy and KL per
    row (ops.categorical_stats) instead of >=10 full passes over [T,B,64,513] / [T,B,16384] tensors;
  * every backward-in-time recursion (6 heads x F fields of V-trace, UPGO, F TD(lambda) returns) is ONE kernel
    launch (ops.return_scan) instead of 2-8 k micro-launches from python loops over T;
  * the ~45 logged scalars leave the device in ONE copy instead of 45 ``.item()`` syncs.
"""
import math
from typing import Dict

import torch

from . import ops

HEADS = ['action_type', 'delay', 'queued', 'selected_units', 'target_unit', 'target_location']
FIELDS = ['winloss', 'build_order', 'built_unit', 'effect', 'upgrade', 'battle']

_DEFAULTS = {  # rl_training/default_reinforcement_loss.yaml
    'loss_weights': {'baseline': {'winloss': 10.0, 'build_order': 0.0, 'built_unit': 0.0, 'effect': 0.0,
                                  'upgrade': 0.0, 'battle': 0.0},
                     'pg': {'winloss': 1.0, 'build_order': 0.0, 'built_unit': 0.0, 'effect': 0.0, 'upgrade': 0.0,
                            'battle': 0.0},
                     'upgo': {'winloss': 1.0}, 'kl': 0.02, 'action_type_kl': 0.1, 'entropy': 0.0001, 'dapo': 0.0},
    'pg_head_weights': {'action_type': 1.0, 'delay': 1.0, 'queued': 1.0, 'selected_units': 0.01, 'target_unit': 1.0,
                        'target_location': 1.0},
    'kl': {'action_type_kl_steps': 2400},
    'dapo': {'dapo_steps': 2400},
    'use_dapo': False,
    'gammas': {'baseline': {'winloss': 1.0, 'build_order': 1.0, 'built_unit': 1.0, 'effect': 1.0, 'upgrade': 1.0,
                            'battle': 0.997}},
}
for _k in ('upgo_head_weights', 'entropy_head_weights', 'kl_head_weights', 'dapo_head_weights'):
    _DEFAULTS[_k] = dict(_DEFAULTS['pg_head_weights'])

# bin/rl_user_config.yaml:58-117 — the values the reference actually trains with
USER_LEARNER_CFG = {
    'loss_weights': {'kl': 0.002, 'action_type_kl': 0.1, 'dapo': 0.1, 'entropy': 0.0001},
    'pg_head_weights': {h: 1.0 for h in HEADS}, 'upgo_head_weights': {h: 1.0 for h in HEADS},
    'entropy_head_weights': {h: 1.0 for h in HEADS}, 'kl_head_weights': {h: 1.0 for h in HEADS},
    'dapo_head_weights': {h: 1.0 for h in HEADS},
    'kl': {'action_type_kl_steps': 5200}, 'dapo': {'dapo_steps': 2400}, 'use_dapo': False,
}


def _merge(a, b):
    out = dict(a)
    for k, v in (b or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


class LazyScalars(dict):
    """dict of logged scalars whose values are still on their way from the device.

    The reference returns ``.item()`` floats (rl_loss.py:40-47 / as_rl_utils); fetching them synchronises host and
    device.  Here all of them travel in one asynchronous copy into pinned memory and turn into floats the first time any
    of them is read (tensor entries such as ``total_loss`` are stored directly and never wait).

    Keys starting with '_' are side-band values (``_total_loss_value``: the float of total_loss; ``_bad_action``: the
    out-of-range-label flag): ``d['_key']`` reads them but they never show up in keys() / items() / iteration, because the
    reference's learner feeds the whole dict to its variable record, which raises on names it has not registered
    (log_helper.py:356-364)."""

    def __init__(self, keys, stacked: torch.Tensor):
        super().__init__()
        self._pending_keys = list(keys)
        self._hidden = {}
        if stacked.is_cuda:
            self._host = torch.empty(stacked.shape, dtype=stacked.dtype, pin_memory=True)
            self._host.copy_(stacked, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()
        else:
            self._host, self._event = stacked, None

    def _settle(self):
        if self._pending_keys is not None:
            if self._event is not None:
                self._event.synchronize()
            keys, self._pending_keys = self._pending_keys, None
            for k, v in zip(keys, self._host.tolist()):
                if k.startswith('_'):
                    self._hidden.setdefault(k, v)
                else:
                    dict.setdefault(self, k, v)

    def __getitem__(self, k):
        if isinstance(k, str) and k.startswith('_'):
            self._settle()
            return self._hidden[k]
        if not dict.__contains__(self, k):
            self._settle()
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        if isinstance(k, str) and k.startswith('_'):
            return False
        return dict.__contains__(self, k) or (self._pending_keys is not None and k in self._pending_keys)

    def get(self, k, default=None):
        return self[k] if k in self else default

    def keys(self):
        self._settle()
        return dict.keys(self)

    def items(self):
        self._settle()
        return dict.items(self)

    def values(self):
        self._settle()
        return dict.values(self)

    def __iter__(self):
        self._settle()
        return dict.__iter__(self)

    def __len__(self):
        self._settle()
        return dict.__len__(self)


class ReinforcementLoss:
    def __init__(self, learner_cfg: dict = None, player_id: str = 'MP0') -> None:
        self.cfg = _merge(_DEFAULTS, learner_cfg if learner_cfg is not None else USER_LEARNER_CFG)
        self.player_id = player_id
        self._load()

    def _load(self):
        c = self.cfg
        self.gammas = c['gammas']
        self.loss_weights = dict(c['loss_weights'])
        self.action_type_kl_steps = c['kl']['action_type_kl_steps']
        self.dapo_steps = c['dapo']['dapo_steps']
        self.use_dapo = bool(c['use_dapo'])
        if 'MP' not in self.player_id:                     # rl_loss.py:22-24
            self.use_dapo = False
            self.loss_weights['dapo'] = 0.0
        self.pg_head_weights = c['pg_head_weights']
        self.upgo_head_weights = c['upgo_head_weights']
        self.entropy_head_weights = c['entropy_head_weights']
        self.kl_head_weights = c['kl_head_weights']
        self.dapo_head_weights = c['dapo_head_weights']
        self.only_update_value = False

    def reset(self, learner_cfg):
        self.cfg = _merge(self.cfg, learner_cfg)
        self._load()

    def compute_loss(self, inputs: dict) -> Dict:
        logits, values = inputs['target_logit'], inputs['value']
        mu, teacher, mask = inputs['action_log_prob'], inputs['teacher_logit'], inputs['mask']
        action, reward, step = inputs['action'], inputs['reward'], inputs['step']
        fields = list(values.keys())
        assert fields and fields[0] == 'winloss', 'winloss baseline must be enabled (UPGO uses it, rl_loss.py:124)'
        T, B = reward['winloss'].shape
        log = {}
        # rl_loss.py:47-49: bootstrap value is zeroed where the last step carried a win/loss reward (visible to caller)
        keep = torch.cat([torch.ones(T, B, device=reward['winloss'].device),
                          (reward['winloss'][-1] == 0).float().unsqueeze(0)], dim=0)
        for f in fields:
            values[f] = values[f] * keep
        su_mask = mask['selected_units_mask']
        su_mask_f = su_mask.float()
        act_mask = {h: (mask['actions_mask'][h].float() if h not in ('action_type', 'delay') else None) for h in HEADS}

        dev = reward['winloss'].device
        if getattr(self, '_flag', None) is None or self._flag.device != dev:
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)     # bit 2: an action label outside its head
        self._flag.zero_()
        lam, ent_rows, kl_rows, rho = {}, {}, {}, {}
        for h in HEADS:                                      # rl_loss.py:63-90 — one fused pass per head
            lp, ent, kl = ops.categorical_stats(logits[h], action[h], teacher[h], flag=self._flag)
            with torch.no_grad():
                lr = lp.detach() - mu[h]
                if h == 'selected_units':
                    lr = (lr * su_mask_f).sum(-1)
                rho[h] = lr.exp().clamp(max=1)
            lam[h] = (lp * su_mask_f).sum(-1) if h == 'selected_units' else lp
            ent_rows[h], kl_rows[h] = ent, kl

        def masked(x, h):
            return x if act_mask[h] is None else x * act_mask[h]

        # as_rl_utils.py:157-312 — every scan in one launch
        gkey = (tuple(fields), keep.device)
        if getattr(self, '_gam_key', None) != gkey:          # cached: building it is a (synchronising) host->device copy
            self._gam = torch.tensor([float(self.gammas['baseline'][f]) for f in fields], device=keep.device)
            self._gam_key = gkey
        gam = self._gam
        vt, up, td = ops.return_scan(torch.stack([reward[f].float() for f in fields]),
                                     torch.stack([values[f] for f in fields]),
                                     torch.stack([rho[h] for h in HEADS]), gam, 0.8)
        total_pg = 0.
        for fi, f in enumerate(fields):                      # as_rl_utils.py:1-28
            tot = 0.
            for hi, h in enumerate(HEADS):
                l = masked(-vt[fi, hi] * lam[h], h)
                if f in ('build_order', 'built_unit', 'effect'):
                    l = l * mask[f + '_mask']
                l = l.mean()
                log['%s/%s' % (f, h)] = l
                tot = tot + l * self.pg_head_weights[h]
            log[f + '/total'] = tot
            total_pg = total_pg + self.loss_weights['pg'][f] * tot
        total_upgo = 0.                                      # as_rl_utils.py:31-49
        for hi, h in enumerate(HEADS):
            l = masked(-up[hi] * lam[h], h).mean()
            log['upgo/' + h] = l
            total_upgo = total_upgo + l * self.upgo_head_weights[h]
        log['upgo/total'] = total_upgo
        total_upgo = total_upgo * self.loss_weights['upgo']['winloss']
        total_critic = 0.                                    # as_rl_utils.py:221-243
        for fi, f in enumerate(fields):
            l = 0.5 * (td[fi] - values[f][:-1]) ** 2
            if f in ('build_order', 'built_unit', 'effect'):
                l = l * mask[f + '_mask']
            l = l.mean()
            total_critic = total_critic + self.loss_weights['baseline'][f] * l
            log[f + '/td'] = l
            log[f + '/reward'] = reward[f].float().mean()
            log[f + '/value'] = values[f].mean()
        log['battle/reward'] = reward['battle'].float().mean()
        total_ent = 0.                                       # as_rl_utils.py:52-72
        for h in HEADS:
            ent = ent_rows[h]
            if h == 'selected_units':
                ent = ent / (1e-9 + torch.log(mask['selected_units_logits_mask'].float().sum(-1) + 1).unsqueeze(-1))
                ent = (ent * su_mask_f).sum(-1) / (su_mask_f.sum(-1) + 1e-9)
            elif h == 'target_unit':
                ent = ent / (1e-9 + torch.log(mask['target_units_logits_mask'].float().sum(-1) + 1))
            else:
                ent = ent / math.log(logits[h].shape[-1])
            ent = masked(ent, h).mean()
            log['entropy/' + h] = ent
            total_ent = total_ent - ent * self.entropy_head_weights[h]
        log['entropy/total'] = total_ent
        total_ent = total_ent * self.loss_weights['entropy']
        total_kl = 0.                                        # as_rl_utils.py:75-103
        at_kl = None
        for h in HEADS:
            kl = kl_rows[h]
            if h == 'selected_units':
                kl = (kl * su_mask_f).sum(-1)
            kl = masked(kl, h)
            if h == 'action_type':
                at_kl = (kl * (step < self.action_type_kl_steps) * mask['cum_action_mask']).mean()
                log['kl/extra_at'] = at_kl
            kl = kl.mean()
            log['kl/' + h] = kl
            total_kl = total_kl + kl * self.kl_head_weights[h]
        log['kl/total'] = total_kl
        total_kl = total_kl * self.loss_weights['kl']
        at_kl = at_kl * self.loss_weights['action_type_kl']
        total_dapo = 0.
        if self.use_dapo:                                    # as_rl_utils.py:105-127: KL(successive || target) per head
            successive = inputs['successive_logit']
            early = (step < self.dapo_steps)
            for h in HEADS:
                kl_d = ops.categorical_stats(logits[h], action[h], successive[h])[2]
                if h == 'selected_units':
                    kl_d = (kl_d * su_mask_f).sum(-1)
                kl_d = (masked(kl_d, h) * early).mean()
                log['battle/' + h] = kl_d                    # the reference logs DAPO under 'battle/*' (as_rl_utils.py:125-126)
                total_dapo = total_dapo + kl_d * self.dapo_head_weights[h]
            log['battle/total'] = total_dapo
            total_dapo = total_dapo * self.loss_weights['dapo']
        if self.only_update_value:
            total = total_critic
        else:
            total = total_pg + total_upgo + total_critic + total_ent + total_kl + at_kl + total_dapo
        # one device->host copy for every logged scalar, and it is not waited for here: the values materialise on first
        # access, so backward can be queued behind the forward pass without draining the launch queue in between
        log['_total_loss_value'] = total         # the float of total_loss rides in the same copy (total_loss itself stays a tensor)
        keys = list(log.keys())
        out = LazyScalars(keys + ['_bad_action'], torch.stack([log[k].detach().float() for k in keys] + [self._flag[0].float()]))
        out['total_loss'] = total
        self._last = out
        return out

    def raise_on_bad_action(self) -> None:
        """torch's Categorical.log_prob raises on an action id outside its head (rl_loss.py:73); the kernels clamp and record
        it, and the record travels with the logged scalars (waiting for it waits only for the loss forward, not backward)."""
        last = getattr(self, '_last', None)
        if last is not None and last['_bad_action'] != 0:
            raise RuntimeError('action id outside its head in the learner batch')
