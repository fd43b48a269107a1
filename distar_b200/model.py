"""``Model`` — drop-in for ``distar.agent.default.model.Model`` (DI-star model/model.py:22-189).

Same constructor, method names, keyword arguments, returned dict keys and ``state_dict`` keys as the reference,
so ``Agent`` (agent.py:127,312,503,725,737), ``RLLearner._train`` (rl_learner.py:102) and ``SLLearner._train``
(sl_learner.py:50) can use it unchanged.  Internals are B200-first:

* all trainable parameters live in ONE contiguous fp32 arena (``flat_param``) with a matching gradient arena
  (``flat_grad``); the per-tensor ``nn.Parameter`` objects the reference API exposes are views into them.  The
  data-parallel exchange is therefore a single NCCL all-reduce of ``flat_grad`` (dist.py) and the optimiser a
  two-kernel pass over the arena (ops.FlatAdam) instead of 394-894 per-tensor calls.
* the compute is ``policy_net.Net`` driving the sm_100a kernels of libdistar_b200.so.
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .constants import SELECTED_UNITS_ACTION_MASK
from .params import init_state_dict
from . import ops
from .policy_net import BASELINE_ATAN, HEADS, MAX_SELECTED_UNITS_NUM, Net, bad_input_message
from .spec import BASELINES, is_trainable, param_specs
from .synth import tree_map
from torch.utils.checkpoint import checkpoint as torch_checkpoint


class _Node(nn.Module):
    """Anonymous container used to rebuild the reference's module tree from dotted state_dict names."""


def _cfg_get(cfg, path, default):
    cur = cfg
    for k in path.split('.'):
        if isinstance(cur, dict) and k in cur:
            cur = cur[k]
        elif hasattr(cur, k) and not isinstance(cur, dict):
            cur = getattr(cur, k)
        else:
            return default
    return cur


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _to_cfg(d):
    return _Cfg({k: _to_cfg(v) for k, v in d.items()}) if isinstance(d, dict) else d


def _async_scalar(t: torch.Tensor):
    if not t.is_cuda:
        v = t.detach().reshape(-1)[0].item()
        return lambda: v
    """Start a device->host copy of a 0-d tensor now; the returned thunk waits for that copy only (an event recorded
    right behind it), not for whatever was queued on the stream afterwards."""
    host = torch.empty(1, dtype=t.dtype, pin_memory=True)
    host.copy_(t.detach().reshape(1), non_blocking=True)
    event = torch.cuda.Event()
    event.record()

    def value():
        event.synchronize()
        return host[0].item()
    return value


class Model(nn.Module):
    def __init__(self, cfg={}, use_value_network=False, temperature=None, seed: Optional[int] = None,
                 gemm_terms: int = 3, sample_rng: str = 'cuda', encoder_chunk: int = 0,
                 checkpoint_encoder: bool = True, keep_chunks: int = 0):
        super().__init__()
        sx = int(_cfg_get(cfg, 'model.spatial_x', 160))
        sy = int(_cfg_get(cfg, 'model.spatial_y', 152))
        self.spatial_x, self.spatial_y = sx, sy
        self.temperature = float(temperature if temperature is not None else _cfg_get(cfg, 'model.temperature', 1.0))
        enabled = list(_cfg_get(cfg, 'model.enable_baselines', BASELINES)) if use_value_network else []
        self.baselines = [b for b in BASELINES if b in enabled]
        # model.py:31-39: the ValueEncoder exists only next to a value network; every baseline then reads
        # [lstm output | value feature | scalar-encoder baseline feature]
        self.use_value_feature = bool(_cfg_get(cfg, 'learner.use_value_feature', False)) and len(self.baselines) > 0
        self.only_update_baseline = bool(_cfg_get(cfg, 'model.only_update_baseline', False))
        self.gemm_terms, self.sample_rng = gemm_terms, sample_rng
        # observation rows are independent in the encoder: process them in chunks (and, while training,
        # recompute each chunk's activations in backward) so B=128 x T=32 fits one GPU's HBM.
        self.encoder_chunk, self.checkpoint_encoder = encoder_chunk, checkpoint_encoder
        self.keep_chunks = keep_chunks          # the first `keep_chunks` chunks keep their activations (HBM permitting)
        # attributes callers read (agent.py:107-108,148)
        self.cfg = _to_cfg({'encoder': {'core_lstm': {'num_layers': 3, 'hidden_size': 384, 'input_size': 1536}},
                            'temperature': self.temperature, 'spatial_x': sx, 'spatial_y': sy,
                            'enable_baselines': self.baselines})
        self._specs = param_specs(sx, sy, self.baselines, self.use_value_feature)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        sd = init_state_dict(seed, sx, sy, self.baselines, perturb=0.0, use_value_feature=self.use_value_feature)
        n_train = sum(int(torch.tensor(s).prod()) if len(s) else 1 for _, s, k in self._specs if is_trainable(k))
        self._offsets = {}
        flat = torch.empty(n_train, dtype=torch.float32)
        off = 0
        for name, shape, kind in self._specs:
            if not is_trainable(kind):
                continue
            n = sd[name].numel()
            off = (off + 3) // 4 * 4          # 16-byte alignment of every tensor inside the arena
            self._offsets[name] = (off, n, tuple(shape))
            off += n
        flat = torch.zeros(off, dtype=torch.float32)
        self._arena_numel = off
        object.__setattr__(self, '_flat_param', flat)
        self._new_grad_arena(flat)
        self._params: Dict[str, nn.Parameter] = {}
        for name, shape, kind in self._specs:
            node = self
            parts = name.split('.')
            for p in parts[:-1]:
                if not hasattr(node, p):
                    node.add_module(p, _Node())
                node = getattr(node, p)
            if is_trainable(kind):
                o, n, shp = self._offsets[name]
                flat[o:o + n].copy_(sd[name].reshape(-1))
                prm = nn.Parameter(flat[o:o + n].view(shp))
            else:
                prm = nn.Parameter(sd[name], requires_grad=False)
            node.register_parameter(parts[-1], prm)
            self._params[name] = prm
        self.policy.action_type_head.race = 'zerg'          # agent.py:148 sets this attribute
        self._bind_grads()
        self._su_mask = torch.tensor(SELECTED_UNITS_ACTION_MASK, dtype=torch.bool)

    # ---------------------------------------------------------------- arena plumbing
    @property
    def flat_param(self) -> torch.Tensor:
        return self._flat_param

    @property
    def flat_grad(self) -> torch.Tensor:
        return self._flat_grad

    _TAIL = 4       # floats behind the gradients that travel in the same all-reduce (slot 0: "this rank's batch was invalid")

    def _new_grad_arena(self, like: torch.Tensor) -> None:
        full = torch.zeros(like.numel() + self._TAIL, dtype=torch.float32, device=like.device)
        object.__setattr__(self, '_flat_grad_full', full)
        object.__setattr__(self, '_flat_grad', full[:like.numel()])

    @property
    def flat_grad_full(self) -> torch.Tensor:
        """flat_grad plus the tail slots: the tensor DistModule.sync_gradients all-reduces."""
        return self._flat_grad_full

    @property
    def grad_tail(self) -> torch.Tensor:
        return self._flat_grad_full[self._arena_numel:]

    def optimizer_layout(self) -> Dict:
        """Where each trainable parameter lives in the arena, by its position in ``parameters()`` - what ops.FlatAdam needs
        to read / write optimizer states in the reference's per-parameter format (torch.optim.Adam over
        ``model.parameters()``, rl_learner.py:73-79; checkpoint_helper.py:124-131,254)."""
        index = {id(p): i for i, p in enumerate(self.parameters())}
        slots = [(index[id(self._params[name])], o, n, shp) for name, (o, n, shp) in self._offsets.items()]
        return {'num_params': len(index), 'slots': sorted(slots)}

    def _bind_grads(self):
        for name, (o, n, shp) in self._offsets.items():
            p = self._params[name]
            p.data = self._flat_param[o:o + n].view(shp)
            p.grad = self._flat_grad[o:o + n].view(shp)

    def _apply(self, fn, recurse=True):
        """Move the arena as a whole (``.cuda()``, ``.to()``, ``.share_memory()``) and re-bind the views."""
        new_flat = fn(self._flat_param)
        object.__setattr__(self, '_flat_param', new_flat)
        self._new_grad_arena(new_flat)
        for name, p in self._params.items():
            if name not in self._offsets:
                p.data = fn(p.data)
        self._bind_grads()
        self._su_mask = fn(self._su_mask)
        ops.invalidate_weight_cache()
        return self

    def zero_grad(self, set_to_none: bool = False):
        self._flat_grad_full.zero_()
        self._bind_grads()

    def load_state_dict(self, state_dict, strict: bool = True):
        missing, unexpected = [], [k for k in state_dict if k not in self._params]
        with torch.no_grad():
            for name, p in self._params.items():
                if name in state_dict:
                    p.data.copy_(state_dict[name].to(p.device))
                else:
                    missing.append(name)
        ops.invalidate_weight_cache()
        if strict and (missing or unexpected):
            raise RuntimeError('load_state_dict: missing %s unexpected %s' % (missing[:5], unexpected[:5]))
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # ---------------------------------------------------------------- compute
    def _net(self) -> Net:
        net = Net(self._params, self.spatial_x, self.spatial_y, self.temperature, self.gemm_terms, self.sample_rng)
        if self.flat_param.is_cuda:
            if getattr(self, '_bad_input_flag', None) is None or self._bad_input_flag.device != self.flat_param.device:
                self._bad_input_flag = torch.zeros(1, dtype=torch.int32, device=self.flat_param.device)
            self._bad_input_flag.zero_()          # one flag per forward pass (a raise may leave queued recomputations behind)
            self._flag_handle = None
            net.bad_input_flag = self._bad_input_flag
        self._last_net = net
        return net

    def raise_on_bad_input(self) -> None:
        """Raise the error the reference raises inside the forward pass for a negative categorical entity id
        (entity_encoder.py:69-72).  compute_logp_action / compute_teacher_logit / sl_train call this before returning;
        rl_learner_forward leaves it to the caller (RLLearner._train checks after backward, before the optimiser step)
        because reading the flag is a host<->device synchronisation."""
        handle = getattr(self, '_flag_handle', None)
        if handle is not None:
            # rl_learner_forward queued an asynchronous copy of the flag right behind the forward pass: waiting for it does
            # not wait for the backward pass the caller has queued since, so the host keeps its lead over the GPU
            self._flag_handle = None
            code = int(handle())
            if code != 0:
                self._bad_input_flag.zero_()
                raise RuntimeError(bad_input_message(code))
            return
        net = getattr(self, '_last_net', None)
        if net is not None:
            net.raise_on_bad_input()

    def _encode(self, net: Net, spatial_info, entity_info, scalar_info, entity_num):
        N = entity_num.shape[0]
        chunk = self.encoder_chunk
        if not chunk or N <= chunk:
            return net.encoder(spatial_info, entity_info, scalar_info, entity_num)

        ckpt = self.checkpoint_encoder and torch.is_grad_enabled()

        state = {'i': 0}

        def entity_fn(en, num):
            # recompute only the entity transformer (~34 MB of saved activations per observation) in backward, and only
            # for the chunks that do not fit the HBM budget
            if ckpt and state['i'] > self.keep_chunks:
                return torch_checkpoint(net.entity_encoder, en, num, use_reentrant=False)
            return net.entity_encoder(en, num)

        outs = []
        # the scalar encoder is tiny and launch-bound: one pass over all rows, sliced per chunk below
        scalar_all = net.scalar_encoder(scalar_info)
        for s0 in range(0, N, chunk):
            state['i'] += 1
            sp, en, num = tree_map(lambda t: t[s0:s0 + chunk], (spatial_info, entity_info, entity_num))
            li, ctx, bf, ee, ms = net.encoder(sp, en, None, num, entity_fn=entity_fn,
                                              scalar_out=tuple(t[s0:s0 + chunk] for t in scalar_all))
            outs.append((li, ctx, bf, ee) + tuple(ms[3:]))        # only the 16x16 skips leave the encoder
        cat = [torch.cat([o[i] for o in outs], dim=0) for i in range(len(outs[0]))]
        return cat[0], cat[1], cat[2], cat[3], [None, None, None] + cat[4:]

    def _value_encode(self, net: Net, vf) -> torch.Tensor:
        """ValueEncoder over all (T+1)*B rows.  Its spatial tower works at full map resolution (~4.5 MB of activations per
        row): it runs in `encoder_chunk`-row chunks and, while training, is recomputed in backward."""
        N = vf['total_unit_count'].shape[0]
        chunk = self.encoder_chunk
        if not chunk or N <= chunk:
            return net.value_encoder(vf)
        ckpt = self.checkpoint_encoder and torch.is_grad_enabled()

        def spatial_fn(*args):
            outs = []
            for s0 in range(0, N, chunk):
                part = tuple(a[s0:s0 + chunk] for a in args)
                outs.append(torch_checkpoint(net.value_encoder_spatial, *part, use_reentrant=False) if ckpt
                            else net.value_encoder_spatial(*part))
            return torch.cat(outs, dim=0)
        return net.value_encoder(vf, spatial_fn=spatial_fn)

    def forward(self, spatial_info, entity_info, scalar_info, entity_num, hidden_state):
        """model.py:46-54."""
        out = self.compute_logp_action(spatial_info, entity_info, scalar_info, entity_num, hidden_state)
        return out['action_info'], out['selected_units_num'], out['hidden_state']

    def compute_logp_action(self, spatial_info, entity_info, scalar_info, entity_num, hidden_state, su_fixed_steps=None,
                            defer_input_check: bool = False, **kwargs):
        """model.py:56-74: encoder -> one LSTM step -> sampling policy -> per-head log-prob of the sample.
        su_fixed_steps / defer_input_check: the host-read-free form serving.InferenceServer captures in a CUDA graph (a fixed
        number of pointer steps; the invalid-input flag is left in ``_bad_input_flag`` for the caller)."""
        net = self._net()
        lstm_input, scalar_context, _bf, entity_embeddings, map_skip = self._encode(
            net, spatial_info, entity_info, scalar_info, entity_num)
        lstm_out, out_state = net.lstm('core_lstm', lstm_input.unsqueeze(0), hidden_state, 3)
        action, su_num, logit, extra = net.policy_sample(lstm_out.squeeze(0), entity_embeddings, map_skip,
                                                        scalar_context, entity_num, self._su_mask, su_fixed_steps)
        logp = {}
        for k, a in action.items():
            logp[k] = torch.log_softmax(logit[k], dim=-1).gather(-1, a.unsqueeze(-1)).squeeze(-1)
        if not defer_input_check:
            net.raise_on_bad_input()
        return {'action_info': action, 'action_logp': logp, 'selected_units_num': su_num, 'entity_num': entity_num,
                'hidden_state': out_state, 'logit': logit, 'extra_units': extra}

    def compute_teacher_logit(self, spatial_info, entity_info, scalar_info, entity_num, hidden_state,
                              selected_units_num, action_info, su_fixed_steps=None, defer_input_check: bool = False, **kwargs):
        """model.py:76-93.  su_fixed_steps / defer_input_check: see compute_logp_action."""
        net = self._net()
        lstm_input, scalar_context, _bf, entity_embeddings, map_skip = self._encode(
            net, spatial_info, entity_info, scalar_info, entity_num)
        lstm_out, out_state = net.lstm('core_lstm', lstm_input.unsqueeze(0), hidden_state, 3)
        _a, su_num, logit = net.policy_train(lstm_out.squeeze(0), entity_embeddings, map_skip, scalar_context,
                                             entity_num, action_info, selected_units_num, su_steps=su_fixed_steps)
        if not defer_input_check:
            net.raise_on_bad_input()
        return {'logit': logit, 'hidden_state': out_state, 'entity_num': entity_num, 'selected_units_num': su_num}

    def rl_learner_forward(self, spatial_info, entity_info, scalar_info, entity_num, hidden_state, action_info,
                           selected_units_num, behaviour_logp, teacher_logit, mask, reward, step, batch_size,
                           unroll_len, **kwargs):
        """model.py:95-168.  Observation rows are time-major [(T+1)*B]; policy runs on the first T*B rows."""
        net = self._net()
        B, T = batch_size, unroll_len
        flat_action = {k: v.flatten(0, 1) for k, v in action_info.items()}
        flat_su_num = selected_units_num.flatten(0, 1)
        # the pointer head loops max(selected_units_num) times: fetch that number with a copy queued BEFORE the encoder so
        # that reading it later waits for nothing (a plain .max().item() there would drain ~200 ms of queued kernels and
        # expose the launch latency of everything after it)
        su_steps = _async_scalar(flat_su_num.max()) if flat_su_num.is_cuda else None
        lstm_input, scalar_context, baseline_feature, entity_embeddings, map_skip = self._encode(
            net, spatial_info, entity_info, scalar_info, entity_num)
        state0 = [(h.view(-1, B, h.shape[-1])[0], c.view(-1, B, c.shape[-1])[0]) for h, c in hidden_state]
        lstm_out, _ = net.lstm('core_lstm', lstm_input.view(-1, B, lstm_input.shape[-1]), state0, 3)
        lstm_out = lstm_out.reshape(-1, lstm_out.shape[-1])
        _a, _n, logits = net.policy_train(lstm_out[:-B], entity_embeddings[:-B], [(m[:-B] if m is not None else None) for m in map_skip],
                                          scalar_context[:-B], entity_num[:-B], flat_action, flat_su_num,
                                          su_steps=su_steps() if su_steps is not None else None)
        critic_input = lstm_out.detach() if self.only_update_baseline else lstm_out
        if self.use_value_feature:                                                  # model.py:141-144
            bf = baseline_feature.detach() if self.only_update_baseline else baseline_feature
            critic_input = torch.cat([critic_input, self._value_encode(net, kwargs['value_feature']), bf], dim=1)
        values = {k: net.value_baseline(k, critic_input).view(T + 1, B) for k in self.baselines}
        logits = {k: v.view(T, B, *v.shape[1:]) for k, v in logits.items()}
        su = logits['selected_units']
        logits['selected_units'] = F.pad(su, (0, 0, 0, MAX_SELECTED_UNITS_NUM - su.shape[2]), 'constant', -1e9)
        self._flag_handle = _async_scalar(self._bad_input_flag) if lstm_out.is_cuda else None
        return {'unroll_len': T, 'batch_size': B, 'selected_units_num': selected_units_num,
                'target_logit': logits, 'value': values, 'action_log_prob': behaviour_logp,
                'teacher_logit': teacher_logit, 'mask': mask, 'action': action_info, 'reward': reward,
                'step': step}

    def sl_train(self, spatial_info, entity_info, scalar_info, entity_num, selected_units_num, traj_lens,
                 hidden_state, action_info, defer_input_check: bool = False, **kwargs):
        """model.py:170-189 (observation rows batch-major [B*T]).  defer_input_check: record an invalid input instead of
        raising here; the caller must call raise_on_bad_input() before the optimiser step (learner.SLLearner does)."""
        net = self._net()
        B = len(traj_lens)
        lstm_input, scalar_context, _bf, entity_embeddings, map_skip = self._encode(
            net, spatial_info, entity_info, scalar_info, entity_num)
        x = lstm_input.view(-1, lstm_input.shape[0] // B, lstm_input.shape[-1]).permute(1, 0, 2)
        lstm_out, out_state = net.lstm('core_lstm', x, hidden_state, 3)
        lstm_out = lstm_out.permute(1, 0, 2).reshape(-1, lstm_out.shape[-1])
        action, su_num, logits = net.policy_train(lstm_out, entity_embeddings, map_skip, scalar_context, entity_num,
                                                  action_info, selected_units_num)
        if defer_input_check and lstm_out.is_cuda:
            self._flag_handle = _async_scalar(self._bad_input_flag)
        else:
            net.raise_on_bad_input()
        return logits, action, out_state
