"""Compact trajectories -> padded learner batch (distar_b200.batch): the expansion must reproduce, bit for
bit, what the reference's host-side collate / padding_entity_info (rl_dataloader.py:45-76,206-245) builds."""
import pytest
import torch

from distar_b200 import ops
from distar_b200.batch import compact_rl_batch, expand_rl_batch, nbytes
from distar_b200.synth import synth_rl_batch, tree_map


def _same(a, b, path=''):
    if isinstance(b, dict):
        assert set(a.keys()) == set(b.keys()), (path, set(a.keys()) ^ set(b.keys()))
        for k in b:
            _same(a[k], b[k], path + '/' + str(k))
    elif isinstance(b, (list, tuple)):
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + '/%d' % i)
    elif torch.is_tensor(b):
        assert a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, b.dtype, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), path
    else:
        assert a == b, path


def _reference_view(batch):
    """the padded batch as the reference's collate builds it: entity slots >= entity_num hold the zero padding of
    padding_entity_info (rl_dataloader.py:213-215; the synthetic generator fills them with values the model masks anyway), and
    the one allowed difference: only the first B rows of the LSTM state travel"""
    B = batch['batch_size']
    ref = dict(batch)
    keep = torch.arange(512).unsqueeze(0) < batch['entity_num'].unsqueeze(1)
    ref['entity_info'] = {k: torch.where(keep, v, torch.zeros((), dtype=v.dtype)) for k, v in batch['entity_info'].items()}
    ref['hidden_state'] = [(h.view(-1, B, h.shape[-1])[0], c.view(-1, B, c.shape[-1])[0]) for h, c in batch['hidden_state']]
    return ref


@pytest.mark.parametrize('B,T,entities', [(3, 4, 'random'), (2, 2, None), (1, 3, 'random')])
def test_expand_inverts_compact_on_cpu(B, T, entities):
    ops.enable_host_logic_testing(True)
    try:
        batch = synth_rl_batch(B, T, seed=B * 10 + T, entity_num=entities, max_su=7)
        if B == 1:                                       # a batch without a single unit selection
            batch = synth_rl_batch(B, T, seed=5, entity_num=entities, max_su=7)
            batch['selected_units_num'].zero_()
            batch['action_info']['selected_units'].zero_()
            batch['behaviour_logp']['selected_units'].fill_(-1e9)
            batch['teacher_logit']['selected_units'].fill_(-1e9)
            batch['mask']['selected_units_mask'].zero_()
        compact = compact_rl_batch(batch)
        _same(expand_rl_batch(compact, 'cpu'), _reference_view(batch))
        assert nbytes(compact) < 0.62 * nbytes(batch)
    finally:
        ops.enable_host_logic_testing(False)


def test_value_feature_travels_through_the_compact_format():
    """learner.use_value_feature: the value_feature entry is already wire-sized (uint8 / bool / int16 fields) and passes through."""
    ops.enable_host_logic_testing(True)
    try:
        batch = synth_rl_batch(2, 3, seed=4, entity_num='random', max_su=5, value_feature=True)
        _same(expand_rl_batch(compact_rl_batch(batch), 'cpu'), _reference_view(batch))
    finally:
        ops.enable_host_logic_testing(False)


@pytest.mark.gpu
@pytest.mark.parametrize('B,T,entities', [(3, 4, 'random'), (4, 8, None)])
def test_expand_on_device_is_bit_exact(B, T, entities):
    batch = synth_rl_batch(B, T, seed=B + T, entity_num=entities, max_su=9)
    compact = compact_rl_batch(batch)
    _same(expand_rl_batch(compact, 'cuda'), _reference_view(batch))


@pytest.mark.gpu
def test_learner_step_from_compact_batch_matches_padded():
    """forward + loss + backward on the device-assembled batch equals the same on the padded batch (loss and the whole gradient
    arena; compared before the optimiser: Adam's first step is sign-like, so fp32 reassociation noise on near-zero gradient
    entries - split-K partial sums meet through the copy engine in arrival order - would be amplified to +-lr)."""
    from distar_b200.model import Model
    from distar_b200.params import init_state_dict
    from distar_b200.rl_loss import ReinforcementLoss
    cfg = {'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}
    batch = synth_rl_batch(2, 3, seed=4, entity_num='random', max_su=6)
    m = Model(cfg, use_value_network=True, seed=0)
    m.load_state_dict(init_state_dict(seed=3))
    m = m.cuda()
    res = []
    for use_compact in (False, True):
        data = expand_rl_batch(compact_rl_batch(batch), 'cuda') if use_compact else tree_map(lambda t: t.cuda(), _reference_view(batch))
        m.zero_grad()
        info = ReinforcementLoss(None, 'MP0').compute_loss(m.rl_learner_forward(**data))
        info['total_loss'].backward()
        res.append((float(info['total_loss']), m.flat_grad.clone()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * max(1.0, abs(res[0][0]))
    rel = (res[0][1] - res[1][1]).norm().item() / res[0][1].norm().item()
    assert rel <= 2e-3, rel                # two runs of the SAME batch differ by this much too (arrival-order sums, ReLU ties)
