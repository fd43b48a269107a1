"""Learner steps — the bodies of ``RLLearner._train`` (DI-star rl_learner.py:82-145) and ``SLLearner._train``
(sl_learner.py:46-76) minus the worker framework (logging, hooks, communication): forward, loss, zero_grad, backward,
gradient sync, clip, Adam — plus what the reference's hooks do to a learner between steps: value pre-training
(rl_learner.py:147-172), checkpoint save / load ({'model', 'optimizer', 'last_iter'}, checkpoint_helper.py:85-140,179-287).

The call sequence and hyper-parameters are the reference's (RL: Adam(betas=(0, 0.99), eps=1e-5) rl_learner.py:73-80,
lr 1e-5 and clip 'pytorch_norm' 1.0 bin/rl_user_config.yaml:40-50; SL: Adam(lr 1e-3, weight_decay 1e-5)
base_learner.py:157-168, clip 'momentum_norm' 1.4 bin/sl_user_config.yaml:31-33); the implementation under each call is
the B200 one (flat arenas, fused kernels, a single NCCL all-reduce).
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from .dist import DistModule, get_world_size
from .model import Model, _async_scalar
from .rl_loss import ReinforcementLoss
from .sl_loss import SupervisedLoss


def _cfg_get(cfg, path, default):
    cur = cfg
    for k in path.split('.'):
        if isinstance(cur, dict) and k in cur:
            cur = cur[k]
        else:
            return default
    return cur


class _LearnerBase:
    """What RL and SL learners share: the arena optimiser, the DP wrapper, checkpoint save / load."""

    def _setup(self, model: Model, lr: float, betas, eps: float, weight_decay: float, clip_type: str, max_norm: float,
               distributed: Optional[bool], group=None):
        self.world = get_world_size(group)          # the learners of THIS player (a league trains several players side by side)
        distributed = self.world > 1 if distributed is None else distributed
        self._model = DistModule(model, group=group) if distributed else model
        self.model = model
        self._optimizer = ops.FlatAdam(model.flat_param, model.flat_grad, lr=lr, betas=betas, eps=eps, max_norm=max_norm,
                                       weight_decay=weight_decay, clip_type=clip_type, layout=model.optimizer_layout(),
                                       owner=model)
        self._distributed = distributed
        self.last_iter = 0

    @property
    def optimizer(self):
        return self._optimizer

    def _sync_clip_step(self, checks: Tuple = ()):
        """Gradient average + clip + Adam, preceded by the invalid-input decision.

        The forward pass only RECORDS invalid inputs (negative entity ids, action labels outside their head) in device
        flags; they are read here, after the whole forward + backward has been queued and before any weight is touched.
        Single process: raise, nothing else happens.  Data parallel: a rank that raised alone would leave the others blocked
        in the NCCL all-reduce forever, and agreeing on the decision through the host would drain the launch queue every
        step.  Instead the flag rides in a spare slot behind the gradients through the SAME all-reduce and the Adam kernel
        skips its update on every rank when the reduced slot is non-zero; the rank that saw the bad batch raises now, the
        others find the reduced slot (an asynchronous copy) when they reach this point in the next iteration."""
        msg = None
        try:
            self.model.raise_on_bad_input()
            for check in checks:
                check()
        except RuntimeError as e:
            msg = str(e)
        if not self._distributed:
            if msg:
                raise RuntimeError(msg)
            return self._optimizer.step(grad_scale=1.0)
        tail = self.model.grad_tail
        if msg:
            tail[0:1].fill_(1.0)
        self._model.sync_gradients()
        gradient = self._optimizer.step(grad_scale=1.0 / self.world, skip_flag=tail)
        previous, self._remote_flag = getattr(self, '_remote_flag', None), _async_scalar(tail[0])
        if msg:
            self._optimizer.t -= 1
            raise RuntimeError(msg)
        if previous is not None and previous() != 0:
            self._optimizer.t -= 1
            raise RuntimeError('invalid input on another rank in the previous iteration: that update was skipped on every rank')
        return gradient

    # ---- checkpoints: the reference's format (checkpoint_helper.py:85-140), so files interchange in both directions
    def state_dict(self) -> Dict:
        return {'model': {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                'optimizer': self._optimizer.state_dict(), 'last_iter': self.last_iter}

    def save_checkpoint(self, path: str) -> None:
        ckpt = self.state_dict()
        ckpt['optimizer'] = {k: (v if k != 'state' else {i: {n: t.cpu() if torch.is_tensor(t) else t for n, t in st.items()}
                                                          for i, st in v.items()}) for k, v in ckpt['optimizer'].items()}
        torch.save(ckpt, path)

    def load_state_dict(self, ckpt: Dict, load_optimizer: bool = True, load_last_iter: bool = True, strict: bool = False):
        """learner_hook.py:145-166 + checkpoint_helper.py:179-287: 'module.'-prefixed keys are accepted, keys of disabled
        value networks are ignored (strict=False), the optimiser / iteration count are optional."""
        model_sd = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in ckpt['model'].items()}
        known = set(self.model.state_dict().keys())
        self.model.load_state_dict({k: v for k, v in model_sd.items() if k in known}, strict=strict)
        if load_optimizer and 'optimizer' in ckpt:
            self._optimizer.load_state_dict(ckpt['optimizer'])
        if load_last_iter and 'last_iter' in ckpt:
            self.last_iter = int(ckpt['last_iter'])

    def load_checkpoint(self, path: str, **kw) -> None:
        self.load_state_dict(torch.load(path, map_location='cpu', weights_only=False), **kw)


class RLLearner(_LearnerBase):
    def __init__(self, model: Model, player_id: str = 'MP0', learner_cfg: dict = None, lr: float = 1e-5,
                 max_norm: float = 1.0, distributed: bool = None, value_pretrain_iters: int = -1,
                 clip_type: str = 'pytorch_norm', group=None):
        self._setup(model, lr, (0.0, 0.99), 1e-5, 0.0, clip_type, max_norm, distributed, group)
        self._loss = ReinforcementLoss(learner_cfg, player_id)
        self._use_dapo = bool(_cfg_get(learner_cfg or {}, 'use_dapo', False))
        self._remain_value_pretrain_iters = int(_cfg_get(learner_cfg or {}, 'value_pretrain_iters', value_pretrain_iters))

    def step_value_pretrain(self) -> None:
        """rl_learner.py:147-172: for the first `value_pretrain_iters` iterations only the baselines learn."""
        if self._remain_value_pretrain_iters > 0:
            self._loss.only_update_value = True
            self.model.only_update_baseline = True
            self._remain_value_pretrain_iters -= 1
        elif self._remain_value_pretrain_iters == 0:
            self._loss.only_update_value = False
            self.model.only_update_baseline = False
            self._remain_value_pretrain_iters -= 1

    def _train(self, data: Dict) -> Dict:
        self.step_value_pretrain()
        data = dict(data)
        data.pop('model_last_iter', None)                       # staleness statistics are logging-only (rl_learner.py:87-101)
        model_output = self._model.rl_learner_forward(**data)
        if self._use_dapo:                                      # rl_learner.py:103-104
            model_output['successive_logit'] = data['successive_logit']
        log_vars = self._loss.compute_loss(model_output)
        loss = log_vars['total_loss']
        self.model.zero_grad()
        loss.backward()
        gradient = self._sync_clip_step((self._loss.raise_on_bad_action,))
        self.last_iter += 1
        log_vars['gradient'] = gradient        # device scalar: norm of the averaged gradient before clipping (no sync here)
        return log_vars

    step = _train   # BASELINE.json calls it rl_learner.step()


class SLLearner(_LearnerBase):
    """sl_learner.py:14-76: LSTM state carried across iterations and reset at episode starts, the first six iterations do
    not update (``ignore_step``), Adam(lr, weight_decay) with warm-up, gradient clip 'momentum_norm'."""

    def __init__(self, model: Model, cfg: dict = None, batch_size: int = 2, lr: float = 1e-3, weight_decay: float = 1e-5,
                 clip_type: str = 'momentum_norm', max_norm: float = 1.4, distributed: bool = None,
                 warm_up_steps: int = 0, ignore_steps: int = 5, group=None):
        cfg = cfg or {}
        lr = float(_cfg_get(cfg, 'learner.learning_rate', lr))
        weight_decay = float(_cfg_get(cfg, 'learner.weight_decay', weight_decay))
        clip_type = _cfg_get(cfg, 'learner.grad_clip.type', clip_type)
        max_norm = float(_cfg_get(cfg, 'learner.grad_clip.threshold', max_norm))
        batch_size = int(_cfg_get(cfg, 'learner.data.batch_size', batch_size))
        self._setup(model, lr, (0.9, 0.999), 1e-8, weight_decay, clip_type, max_norm, distributed, group)
        self._loss = SupervisedLoss(cfg)
        dev = model.flat_param.device
        layers, hidden = model.cfg.encoder.core_lstm.num_layers, model.cfg.encoder.core_lstm.hidden_size
        self.hidden_state: List[Tuple[torch.Tensor, torch.Tensor]] = [
            (torch.zeros(batch_size, hidden, device=dev), torch.zeros(batch_size, hidden, device=dev)) for _ in range(layers)]
        self.ignore_step, self._ignore_steps, self._updates = 0, ignore_steps, 0
        self._base_lr, self._warm_up_steps = lr, int(_cfg_get(cfg, 'learner.warm_up_steps', warm_up_steps)) \
            if _cfg_get(cfg, 'learner.use_warmup', warm_up_steps > 0) else 0

    def reset_hidden_state(self, new_episodes) -> None:
        """sl_learner.py:32-36."""
        keep = torch.ones(self.hidden_state[0][0].shape[0], 1, device=self.hidden_state[0][0].device)
        keep[new_episodes] = 0
        self.hidden_state = [(h.detach() * keep, c.detach() * keep) for h, c in self.hidden_state]

    def _train(self, data: Dict) -> Dict:
        data = dict(data)
        new_episodes = data.pop('new_episodes', [])
        self.reset_hidden_state(new_episodes)
        logits, infer_action_info, hidden_state = self._model.sl_train(**data, hidden_state=self.hidden_state,
                                                                       defer_input_check=True)
        log_vars = self._loss.compute_loss(logits, data['action_info'], data['action_mask'], data['selected_units_num'],
                                           data['entity_num'], infer_action_info)
        loss = log_vars['total_loss']
        if self.ignore_step > self._ignore_steps:
            self.model.zero_grad()
            loss.backward()
            if self._warm_up_steps:
                # GradualWarmupScheduler(multiplier=1) (lr_scheduler_util.py:31-40), stepped after every update
                # (sl_learner.py:71-72): the k-th update runs at base_lr * k / warm_up_steps, from 0 up to the base rate
                self._optimizer.lr = self._base_lr * min(1.0, self._updates / self._warm_up_steps)
            gradient = self._sync_clip_step()
            self._updates += 1
        else:
            self.model.raise_on_bad_input()        # no update this iteration, but an invalid batch is still an error
            gradient = 0.
        self.ignore_step += 1
        self.last_iter += 1
        self.hidden_state = [(h.detach(), c.detach()) for h, c in hidden_state]
        log_vars['gradient'] = gradient
        return log_vars

    step = _train
