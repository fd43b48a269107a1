"""Data-parallel plumbing — drop-in for ``distar/ctools/utils/dist_helper.py`` (dist_init, DistModule, allreduce,
broadcast; reference lines 259-439).

The reference averages gradients with one blocking ``all_reduce`` + ``div_`` PER PARAMETER TENSOR (394-894 NCCL
calls per step, dist_helper.py:421-431).  Here the gradients already live in one contiguous arena
(Model.flat_grad), so ``sync_gradients`` is ONE NCCL all-reduce over NVLink/NVSwitch and the division by the
world size is folded into the optimiser kernel (ops.FlatAdam grad_scale).  Inference needs no collective.
"""
import os

import torch
import torch.distributed as dist


def dist_init(backend: str = 'nccl', init_method: str = None, rank: int = None, world_size: int = None):
    """dist_helper.py:321-344: env:// (torchrun) or explicit tcp:// rendezvous; one process per GPU."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    rank = int(os.environ.get('RANK', 0)) if rank is None else rank
    world_size = int(os.environ.get('WORLD_SIZE', 1)) if world_size is None else world_size
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1))))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    dist.init_process_group(backend=backend, init_method=init_method or 'env://', rank=rank, world_size=world_size)
    return rank, world_size


def get_rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def get_world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def allreduce(x: torch.Tensor, reduce: bool = True):
    """dist_helper.py:272-289 (SUM, optionally divided by world)."""
    if get_world_size() > 1:
        dist.all_reduce(x)
        if reduce:
            x.div_(get_world_size())
    return x


def broadcast(x: torch.Tensor, src: int = 0):
    if get_world_size() > 1:
        dist.broadcast(x, src)
    return x


class DistModule(torch.nn.Module):
    """dist_helper.py:369-439.  Exposes forward / state_dict / named_parameters / sl_train / rl_learner_forward of
    the wrapped model, ``.module``, ``sync_gradients()`` and ``broadcast_params()``."""

    def __init__(self, module, sync: bool = True, group=None):
        """group: the process group of THIS player's learners.  A league run trains several players at once, each on its own
        set of GPUs (rl_train.py:26-51 starts one learner job per player id); their gradient exchanges are independent
        communicators (BASELINE configs[4]); None = the default (world) group."""
        super().__init__()
        self.module = module
        self.group = group
        for name in ('compute_logp_action', 'compute_teacher_logit', 'rl_learner_forward', 'sl_train'):
            if hasattr(module, name):
                setattr(self, name, getattr(module, name))
        self.sync = sync
        self.broadcast_params()

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self.module.load_state_dict(*a, **k)

    def named_parameters(self, *a, **k):
        return self.module.named_parameters(*a, **k)

    def parameters(self, *a, **k):
        return self.module.parameters(*a, **k)

    def sync_gradients(self):
        """ONE all-reduce(SUM) of the flat gradient arena; the 1/world average is applied by the optimiser."""
        if self.sync and get_world_size(self.group) > 1:
            # the tail slots behind the gradients ride along (learner.py: the "a rank saw an invalid batch" flag)
            dist.all_reduce(getattr(self.module, 'flat_grad_full', self.module.flat_grad), group=self.group)

    def broadcast_params(self):
        from . import ops
        if get_world_size(self.group) > 1:
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast(self.module.flat_param, src, group=self.group)
            for name, p in self.module.named_parameters():
                if not p.requires_grad:
                    dist.broadcast(p.data, src, group=self.group)
        ops.invalidate_weight_cache()
