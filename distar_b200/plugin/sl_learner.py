"""distar/agent/b200/sl_learner.py — ``SLLearner`` (distar/agent/default/sl_learner.py) with the hot path swapped the same way
as rl_learner.py: model, loss, optimiser (+ the 'momentum_norm' clip of bin/sl_user_config.yaml:31-33) and DP wrapper."""
import torch

from distar.ctools.worker.learner.base_learner import BaseLearner
from distar.ctools.torch_utils.grad_clip import build_grad_clip

from distar_b200 import ops
from distar_b200.dist import DistModule
from distar_b200.model import Model
from distar_b200.sl_loss import SupervisedLoss


class SLLearner(BaseLearner):
    def _init_model(self):
        self._setup_model()
        if self._use_cuda:
            self._model = self._model.to(device=self._device)
        if self.use_distributed:
            self._model = DistModule(self._model)

    def _setup_model(self):                                                   # sl_learner.py:15-24
        self._model = Model(self._whole_cfg, temperature=1.0)
        self._grad_clip = build_grad_clip(self._whole_cfg.learner.grad_clip)
        self.num_layers = self._model.cfg.encoder.core_lstm.num_layers
        self.hidden_size = self._model.cfg.encoder.core_lstm.hidden_size
        zero = torch.zeros(self._whole_cfg.learner.data.batch_size, self.hidden_size)
        if self._whole_cfg.learner.use_cuda and torch.cuda.is_available():
            zero = zero.cuda()
        self.hidden_state = [(zero, zero) for _ in range(self.num_layers)]
        self.ignore_step = 0

    def _setup_loss(self):
        self._loss = SupervisedLoss(self._whole_cfg)

    def _setup_optimizer(self):                                               # base_learner.py:157-181, Adam -> FlatAdam
        cfg = self._whole_cfg.learner
        m = self._model.module if hasattr(self._model, 'module') else self._model
        self._optimizer = ops.FlatAdam(m.flat_param, m.flat_grad, lr=cfg.learning_rate, betas=(0.9, 0.999), eps=1e-8,
                                       weight_decay=cfg.weight_decay, max_norm=cfg.grad_clip.get('threshold', 1.4),
                                       clip_type=cfg.grad_clip.get('type', 'none'), layout=m.optimizer_layout(), owner=m)
        # the schedulers are the reference's own (base_learner.py:168-181), acting on FlatAdam's param group
        decay, interval = cfg.get('lr_decay', 1.), int(cfg.get('lr_decay_interval', 1e20))
        if cfg.get('use_warmup', False):
            from distar.ctools.torch_utils.lr_scheduler_util import GradualWarmupScheduler
            decay, interval = cfg.get('lr_decay', 0.9), int(cfg.get('lr_decay_interval', 10000))
            self._after_lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(
                self._optimizer, milestones=list(range(0, interval * 40, interval))[1:], gamma=decay)
            self._lr_scheduler = GradualWarmupScheduler(optimizer=self._optimizer, multiplier=cfg.get('multiplier', 1),
                                                        total_epoch=cfg.get('warm_up_steps', 10000),
                                                        after_scheduler=self._after_lr_scheduler)
        else:
            self._lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(
                self._optimizer, milestones=list(range(0, interval * 20, interval))[1:], gamma=decay)

    def _setup_dataloader(self):                                              # sl_learner.py:40-44
        if self._whole_cfg.learner.job_type == 'train':
            from distar.agent.default.sl_training.sl_dataloader import SLDataloader     # the reference's replay-decoding loader
            self._dataloader = SLDataloader(self._whole_cfg)
            return
        # any other job type: seeded synthetic batches (the role of the reference's FakeDataloader)
        from distar_b200.synth import synth_sl_batch, tree_map
        cfg = self._whole_cfg.learner.data
        batch = tree_map(lambda t: t.to(self._device), synth_sl_batch(cfg.batch_size, cfg.trajectory_length, seed=0))
        self._dataloader = iter(lambda: dict(batch), None)

    def reset_hidden_state(self, new_episodes):                               # sl_learner.py:32-36
        keep = torch.ones(self.hidden_state[0][0].shape[0], 1, device=self.hidden_state[0][0].device)
        keep[new_episodes] = 0
        self.hidden_state = [(h.detach() * keep, c.detach() * keep) for h, c in self.hidden_state]

    def _train(self, data):                                                   # sl_learner.py:46-76
        m = self._model.module if hasattr(self._model, 'module') else self._model
        with self._timer:
            data = dict(data)
            self.reset_hidden_state(data.pop('new_episodes', []))
            logits, infer_action_info, hidden_state = self._model.sl_train(**data, hidden_state=self.hidden_state)
            log_vars = self._loss.compute_loss(logits, data['action_info'], data['action_mask'], data['selected_units_num'],
                                               data['entity_num'], infer_action_info)
            loss = log_vars['total_loss']
        self._log_buffer['forward_time'] = self._timer.value
        with self._timer:
            if self.ignore_step > 5:
                m.zero_grad()
                loss.backward()
                if self._use_distributed:
                    self._model.sync_gradients()
                gradient = float(self._optimizer.step(grad_scale=1.0 / self._world_size))
                self._lr_scheduler.step()
            else:
                gradient = 0.
            self.ignore_step += 1
        self.hidden_state = [(h.detach(), c.detach()) for h, c in hidden_state]
        self._log_buffer['gradient'] = gradient
        self._log_buffer['backward_time'] = self._timer.value
        self._log_buffer.update({k: (v.item() if torch.is_tensor(v) and v.numel() == 1 and k != 'total_loss' else v)
                                 for k, v in log_vars.items()})
