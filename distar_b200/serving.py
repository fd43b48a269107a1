"""Actor-side serving of the policy and learner -> actor weight publication.

Reference: the actor runs ``Model.compute_logp_action`` (and the teacher's / successive model's ``compute_teacher_logit``)
once per game step for a batch of environments, from a process that hands observations over through shared memory and
sleep-polled signal flags (agent.py:298-319,715-739; actor.py:268-299); the learner publishes weights by copying the
state_dict to CPU shared memory, ``torch.save`` + lz4 and a TCP send every few iterations (learner_comm.py:72-99).

B200-first replacements, same roles:

* ``InferenceServer``: the fixed-shape ``compute_logp_action`` -> ``compute_teacher_logit`` pair captured ONCE into a CUDA
  graph over static device buffers.  A request is: copy the observations into the static buffers (pinned host -> device),
  replay, copy the results back.  The ~800 kernel launches of a batch-32 call (of which the pointer network alone is 64 fixed
  steps) cost one graph launch instead of ~20 ms of Python + launch overhead.  Inference needs no collective: one server per
  GPU.
* ``WeightPublisher`` / ``WeightSubscriber``: the policy weights live in ONE contiguous arena, so a publication is one
  device -> pinned-host copy of the arena segments an actor needs (value networks excluded, as learner_comm.py:74 does) with a
  version counter, and a subscription is one host -> device copy into the actor model's arena followed by an IN-PLACE refresh
  of the derived weight forms (bf16 pairs, conv matrices) the captured graph reads — same addresses, so the graph stays valid.

Transport between machines (the reference's TCP / file system adapters) is outside the hot path: the published buffer is a
plain pinned tensor that any transport can ship.
"""
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .model import Model
from .policy_net import HEADS, bad_input_message
from .synth import tree_map


def _copy_tree(dst, src):
    if isinstance(src, dict):
        for k in src:
            _copy_tree(dst[k], src[k])
    elif isinstance(src, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_tree(d, s)
    elif torch.is_tensor(src):
        dst.copy_(src, non_blocking=True)


class InferenceServer:
    """CUDA-graph replay of ``model.compute_logp_action`` (and optionally ``teacher.compute_teacher_logit`` on the sampled
    action) for a fixed batch size.  ``example_obs``: one observation batch of the serving shape (dict with spatial_info,
    entity_info, scalar_info, entity_num, hidden_state) used to size the static buffers.

    su_steps: pointer-network steps captured (the reference stops when every row has picked its end token, at most 64;
    a captured graph cannot branch on device data, so it always runs `su_steps` and the outputs are trimmed to the reference's
    length afterwards — rows that ended earlier are unaffected by the extra steps)."""

    def __init__(self, model: Model, example_obs: Dict, teacher: Optional[Model] = None, su_steps: int = 64,
                 warmup: int = 2):
        assert model.flat_param.is_cuda, 'InferenceServer needs a CUDA model (no CPU fallback)'
        self.model, self.teacher, self.su_steps = model, teacher, su_steps
        dev = model.flat_param.device
        self.device = dev
        self.static_in = tree_map(lambda t: t.to(dev).clone(), example_obs)
        self.teacher_hidden = None
        if teacher is not None:
            self.teacher_hidden = [(h.clone(), c.clone()) for h, c in self.static_in['hidden_state']]
        self.graph = None
        self.static_out = None
        self.replays = 0
        self._capture(warmup)

    def _forward(self):
        obs = self.static_in
        out = self.model.compute_logp_action(**obs, su_fixed_steps=self.su_steps, defer_input_check=True)
        res = {'action_info': out['action_info'], 'action_logp': out['action_logp'],
               'selected_units_num': out['selected_units_num'], 'logit': out['logit'], 'hidden_state': out['hidden_state'],
               'bad_input': self.model._bad_input_flag}
        if self.teacher is not None:
            t_in = dict(obs, hidden_state=self.teacher_hidden)
            t = self.teacher.compute_teacher_logit(**t_in, selected_units_num=out['selected_units_num'],
                                                   action_info=out['action_info'], su_fixed_steps=self.su_steps,
                                                   defer_input_check=True)
            res['teacher_logit'] = t['logit']
            res['teacher_hidden_state'] = t['hidden_state']
        return res

    def _capture(self, warmup: int):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):           # populates the weight-form caches and the allocator before capture
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self._forward()
        torch.cuda.synchronize()

    @torch.no_grad()
    def infer(self, obs: Dict, teacher_hidden: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None) -> Dict:
        """obs: tensors on the host (ideally pinned) or device with the captured shapes.  Returns device tensors trimmed to
        the reference's shapes: selected-units outputs have max(selected_units_num) steps (at least 1)."""
        _copy_tree(self.static_in, obs)
        if teacher_hidden is not None and self.teacher_hidden is not None:
            _copy_tree(self.teacher_hidden, teacher_hidden)
        self.graph.replay()
        self.replays += 1
        out = self.static_out
        code, steps = torch.stack([out['bad_input'][0].to(torch.int64), out['selected_units_num'].max()]).tolist()   # ONE host read
        if code:
            raise RuntimeError(bad_input_message(int(code)))
        steps = max(1, min(int(steps), self.su_steps))
        res = {'action_info': {k: v.clone() for k, v in out['action_info'].items()},
               'action_logp': {k: v.clone() for k, v in out['action_logp'].items()},
               'selected_units_num': out['selected_units_num'].clone(),
               'logit': {k: v.clone() for k, v in out['logit'].items()},
               'hidden_state': [(h.clone(), c.clone()) for h, c in out['hidden_state']]}
        for d in (res['action_info'], res['action_logp'], res['logit']):
            d['selected_units'] = d['selected_units'][:, :steps].contiguous()
        if 'teacher_logit' in out:
            res['teacher_logit'] = {k: v.clone() for k, v in out['teacher_logit'].items()}
            res['teacher_logit']['selected_units'] = res['teacher_logit']['selected_units'][:, :steps].contiguous()
            res['teacher_hidden_state'] = [(h.clone(), c.clone()) for h, c in out['teacher_hidden_state']]
        return res

    def weights_updated(self, which: Optional[Model] = None) -> int:
        """Call after new weights were written into a served model's arena: the derived weight forms the captured launches
        read (bf16 pairs, conv GEMM matrices, padded heads) are recomputed into the SAME buffers."""
        n = 0
        for m in ([which] if which is not None else [self.model, self.teacher]):
            if m is not None:
                ops.WEIGHT_EPOCH[0] += 1
                n += ops.refresh_weight_cache(m.parameters())
        return n


class WeightPublisher:
    """Learner side: ``publish()`` snapshots the policy part of the arena into pinned host memory (asynchronously, on a side
    stream) and bumps ``version`` (learner_comm.py:72-99 sends every `send_model_freq` iterations).  ``names`` are the
    parameters an actor holds: everything but the value networks (learner_comm.py:74, actor.py:71-73)."""

    def __init__(self, model: Model):
        self.model = model
        # learner_comm.py:74: actors never see the critic ('value_networks' / 'value_encoder' keys stay on the learner)
        self.names = [n for n in model._offsets if not n.startswith(('value_networks', 'value_encoder'))]
        segs = sorted((model._offsets[n][0], model._offsets[n][1], n) for n in self.names)
        # coalesce neighbouring slots (16-byte alignment gaps are copied along) into a few contiguous ranges
        self.ranges: List[List[int]] = []
        for off, n, _ in segs:
            if self.ranges and off - (self.ranges[-1][0] + self.ranges[-1][1]) < 4:
                self.ranges[-1][1] = off + n - self.ranges[-1][0]
            else:
                self.ranges.append([off, n])
        self.layout = {n: (model._offsets[n][0], model._offsets[n][1]) for n in self.names}
        total = sum(n for _, n in self.ranges)
        pin = model.flat_param.is_cuda
        self.buffer = torch.empty(total, dtype=torch.float32, pin_memory=pin)
        self.version = 0
        self._stream = torch.cuda.Stream(device=model.flat_param.device) if pin else None
        self._event = None

    def publish(self) -> int:
        src = self.model.flat_param
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())       # after the optimiser step that produced these weights
            with torch.cuda.stream(self._stream):
                pos = 0
                for off, n in self.ranges:
                    self.buffer[pos:pos + n].copy_(src[off:off + n], non_blocking=True)
                    pos += n
                self._event = torch.cuda.Event()
                self._event.record(self._stream)
        else:
            pos = 0
            for off, n in self.ranges:
                self.buffer[pos:pos + n].copy_(src[off:off + n])
                pos += n
        self.version += 1
        return self.version

    def wait(self) -> torch.Tensor:
        """The published snapshot, complete (what a transport would ship)."""
        if self._event is not None:
            self._event.synchronize()
        return self.buffer

    def segment(self, name: str) -> Tuple[int, int]:
        """(offset, numel) of parameter `name` inside the published buffer."""
        off, n = self.layout[name]
        pos = 0
        for roff, rn in self.ranges:
            if roff <= off < roff + rn:
                return pos + off - roff, n
            pos += rn
        raise KeyError(name)


class WeightSubscriber:
    """Actor side: ``update(buffer, publisher_layout)`` writes a published snapshot into the actor model's own arena (its
    layout differs: no value networks) and refreshes the derived weight forms in place."""

    def __init__(self, model: Model, server: Optional[InferenceServer] = None):
        self.model, self.server, self.version = model, server, 0

    def update(self, publisher: WeightPublisher, version: Optional[int] = None) -> None:
        buf = publisher.wait()
        dst = self.model.flat_param
        # contiguous runs of parameters that are neighbours in BOTH arenas travel as one copy
        runs: List[List[int]] = []
        for name, (doff, n, _shape) in self.model._offsets.items():
            if name not in publisher.layout:
                continue
            soff, sn = publisher.segment(name)
            assert sn == n, name
            if runs and runs[-1][0] + runs[-1][2] == soff and runs[-1][1] + runs[-1][2] == doff:
                runs[-1][2] += n
            elif runs and 0 < soff - (runs[-1][0] + runs[-1][2]) < 4 and soff - (runs[-1][0] + runs[-1][2]) == doff - (runs[-1][1] + runs[-1][2]):
                runs[-1][2] = soff + n - runs[-1][0]
            else:
                runs.append([soff, doff, n])
        with torch.no_grad():
            for soff, doff, n in runs:
                dst[doff:doff + n].copy_(buf[soff:soff + n], non_blocking=True)
        self.version = version if version is not None else publisher.version
        if self.server is not None:
            self.server.weights_updated(self.model)
        else:
            ops.invalidate_weight_cache()
