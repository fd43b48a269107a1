"""distar/agent/b200/rl_learner.py — ``RLLearner`` under the reference's own worker framework (BaseLearner: hooks, logging,
checkpoints, league communication stay the reference's), with the hot path swapped:

  _setup_model      distar_b200 Model (one parameter arena)                                   rl_learner.py:55-59
  _setup_loss       distar_b200 ReinforcementLoss (fused per-head statistics + return scans)   rl_learner.py:61-62
  _setup_optimizer  ops.FlatAdam: clip + Adam over the arena in two kernels; a torch Optimizer, so the MultiStepLR and the
                    checkpoint hooks of the framework work on it unchanged                      rl_learner.py:73-80
  _init_model       the DP wrapper is distar_b200.dist.DistModule (ONE NCCL all-reduce)         base_learner.py:96-101
  _train            the reference's step (rl_learner.py:82-145) with gradient sync, clip and Adam fused into
                    sync_gradients() + optimizer.step()
"""
import torch

from distar.agent.default.rl_learner import RLLearner as _Base
from distar.ctools.torch_utils.grad_clip import build_grad_clip

from distar_b200 import ops
from distar_b200.dist import DistModule
from distar_b200.model import Model
from distar_b200.rl_loss import ReinforcementLoss


class _SyntheticDataloader:
    """job_type != 'train' (no league to pull trajectories from): seeded synthetic learner batches, the role of the
    ``FakeDataloader`` the stock class refers to (rl_learner.py:192-196)."""

    def __init__(self, batch_size, unroll_len, device, value_feature=False):
        from distar_b200.synth import synth_rl_batch, tree_map
        self._batch = tree_map(lambda t: t.to(device), synth_rl_batch(batch_size, unroll_len, seed=0, value_feature=value_feature))

    def __iter__(self):
        return self

    def __next__(self):
        data = dict(self._batch)
        data['model_last_iter'] = torch.zeros(1)
        return data


class RLLearner(_Base):
    def _init_model(self):
        self._setup_model()
        if self._use_cuda:
            self._model = self._model.to(device=self._device)
        if self.use_distributed:
            self._model = DistModule(self._model)
        self._grad_clip = build_grad_clip(self._whole_cfg.learner.grad_clip)     # kept for the hooks that read it

    def _setup_model(self):
        self._model = Model(self._whole_cfg, use_value_network=True)

    def _setup_loss(self):
        self._loss = ReinforcementLoss(self._whole_cfg.learner, self._whole_cfg.learner.player_id)

    def _setup_optimizer(self):
        m = self.model.module if hasattr(self.model, 'module') else self.model
        clip = self._whole_cfg.learner.grad_clip
        self._optimizer = ops.FlatAdam(m.flat_param, m.flat_grad, lr=self._whole_cfg.learner.learning_rate, betas=(0.0, 0.99),
                                       eps=1e-5, max_norm=clip.get('threshold', 1.0), clip_type=clip.get('type', 'pytorch_norm'),
                                       layout=m.optimizer_layout(), owner=m)
        self._lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(self._optimizer, milestones=[], gamma=1)

    def _setup_dataloader(self):
        if self._job_type != 'train':
            self._dataloader = _SyntheticDataloader(self._whole_cfg.learner.data.batch_size,
                                                    self._whole_cfg.actor.get('traj_len', 2), self._device,
                                                    bool(self._whole_cfg.learner.get('use_value_feature', False)))

    def _get_iter_data(self):
        return next(self._dataloader)

    def _reset_value(self):                                                   # rl_learner.py:225-228
        """League reset of the critic: fresh value networks (and ValueEncoder), policy untouched."""
        fresh = Model(self._whole_cfg, use_value_network=True)
        value_state_dict = {k: v for k, v in fresh.state_dict().items() if 'value' in k or 'auxiliary' in k}
        self.model.load_state_dict(value_state_dict, strict=False)

    @property
    def model(self):
        return self._model

    def _train(self, data):
        m = self._model.module if hasattr(self._model, 'module') else self._model
        with self._timer:
            self.step_value_pretrain()
            data = dict(data)
            model_last_iter = data.pop('model_last_iter', None)
            staleness = staleness_std = staleness_max = 0
            if self._remain_value_pretrain_iters <= 0 and model_last_iter is not None:      # rl_learner.py:87-101
                diff = self.last_iter.val - model_last_iter
                if diff.shape[0] == 1:
                    staleness = staleness_max = diff.item()
                else:
                    staleness_std, staleness = (v.item() for v in torch.std_mean(diff))
                    staleness_max = torch.max(diff).item()
            model_output = self._model.rl_learner_forward(**data)
            if self._whole_cfg.learner.get('use_dapo', False):
                model_output['successive_logit'] = data['successive_logit']
            log_vars = self._loss.compute_loss(model_output)
            log_vars['entropy/reward'] = staleness          # (the reference logs staleness under these names, :106-108)
            log_vars['entropy/value'] = staleness_std
            log_vars['entropy/td'] = staleness_max
            loss = log_vars['total_loss']
        self._log_buffer['forward_time'] = self._timer.value
        with self._timer:
            m.zero_grad()
            loss.backward()
            m.raise_on_bad_input()
            self._loss.raise_on_bad_action()
            if self._use_distributed:
                self._model.sync_gradients()
            scale = 1.0 / self._world_size                 # the arena holds the SUM over ranks; the step folds the average in
            save = getattr(self, '_save_grad', False) and self._last_iter.val % self.save_log_freq == 0
            if save:                                                               # rl_learner.py:118-124
                for k, param in m.named_parameters():
                    if param.grad is not None and param.requires_grad:
                        self.grad_tb_logger.add_scalar(k, torch.norm(param.grad).item() * scale, global_step=self._last_iter.val)
                        self.model_tb_logger.add_scalar(k, torch.norm(param.data).item(), global_step=self._last_iter.val)
            gradient = self._optimizer.step(grad_scale=scale)                       # clip_grad_norm_ + Adam, fused
            if save:                                                               # rl_learner.py:126-130 (gradients after the clip)
                max_norm = self._optimizer.max_norm
                coef = min(1.0, max_norm / (float(gradient) + 1e-6)) if max_norm is not None else 1.0
                for k, param in m.named_parameters():
                    if param.grad is not None and param.requires_grad:
                        self.clip_grad_tb_logger.add_scalar(k, torch.norm(param.grad).item() * scale * coef,
                                                            global_step=self._last_iter.val)
        self._log_buffer['gradient'] = float(gradient)
        self._log_buffer['backward_time'] = self._timer.value
        self._log_buffer.update(log_vars)
        for flag, fn in (('_update_config_flag', 'update_config'), ('_reset_value_flag', 'reset_value'),
                         ('_reset_comm_setting_flag', 'reset_comm_setting')):
            if getattr(self, flag, False):
                getattr(self, fn)()
                setattr(self, flag, False)
