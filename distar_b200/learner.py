"""One optimisation step of the RL learner — the body of ``RLLearner._train`` (DI-star rl_learner.py:82-145)
minus logging/hooks: forward, loss, zero_grad, backward, gradient sync, clip, Adam.

The call sequence and hyper-parameters are the reference's (Adam(betas=(0, 0.99), eps=1e-5) rl_learner.py:73-80,
lr 1e-5 and clip 'pytorch_norm' 1.0 bin/rl_user_config.yaml:40-50); the implementation under each call is the
B200 one (flat arenas, fused kernels, a single NCCL all-reduce).
"""
from typing import Dict

import torch

from . import ops
from .dist import DistModule, get_world_size
from .model import Model
from .rl_loss import ReinforcementLoss


class RLLearner:
    def __init__(self, model: Model, player_id: str = 'MP0', learner_cfg: dict = None, lr: float = 1e-5,
                 max_norm: float = 1.0, distributed: bool = None):
        self.world = get_world_size()
        distributed = self.world > 1 if distributed is None else distributed
        self._model = DistModule(model) if distributed else model
        self.model = model
        self._loss = ReinforcementLoss(learner_cfg, player_id)
        self._optimizer = ops.FlatAdam(model.flat_param, model.flat_grad, lr=lr, betas=(0.0, 0.99), eps=1e-5,
                                       max_norm=max_norm)
        self._distributed = distributed
        self.last_iter = 0

    def _train(self, data: Dict) -> Dict:
        model_output = self._model.rl_learner_forward(**data)
        log_vars = self._loss.compute_loss(model_output)
        loss = log_vars['total_loss']
        self.model.zero_grad()
        loss.backward()
        # the forward pass only records invalid inputs (negative entity ids) in a device flag; raise here, after the whole
        # forward + backward has been queued and before any weight is touched
        self.model.raise_on_bad_input()
        if self._distributed:
            self._model.sync_gradients()
        gradient = self._optimizer.step(grad_scale=1.0 / self.world)
        self.last_iter += 1
        log_vars['gradient'] = gradient        # device scalar (no sync here)
        return log_vars

    step = _train   # BASELINE.json calls it rl_learner.step()
