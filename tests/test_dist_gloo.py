"""CPU, world_size 2 (gloo): the flat-arena gradient exchange of dist.DistModule equals the per-tensor average the
reference computes (ctools/utils/dist_helper.py:421-431)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from distar_b200 import ops
    from distar_b200.dist import DistModule, dist_init, get_world_size
    from distar_b200.model import Model
    dist_init('gloo')
    ops.enable_host_logic_testing(True)
    torch.manual_seed(rank)      # different init per rank: broadcast_params must make them equal
    m = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True)
    dm = DistModule(m)
    ref = m.flat_param.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, m.flat_param))
    g = torch.Generator().manual_seed(100 + rank)
    m.flat_grad.copy_(torch.randn(m.flat_grad.shape, generator=g))
    mine = m.flat_grad.clone()
    dm.sync_gradients()
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    expect = sum(both)
    ok = bool(torch.allclose(m.flat_grad, expect)) and get_world_size() == world
    # the optimiser folds 1/world (grad_scale): parameters stay identical across ranks after a step
    opt = ops.FlatAdam(m.flat_param, m.flat_grad, lr=1e-3, max_norm=1.0)
    opt.step(grad_scale=1.0 / world)
    after = m.flat_param.clone()
    dist.broadcast(after, 0)
    ok = ok and bool(torch.allclose(after, m.flat_param))
    out.put((rank, same, ok))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(same and ok for _, same, ok in res), res
