"""Pins the oracle against the committed fixtures dumped from the real reference (runs anywhere, CPU)."""
import os

import pytest
import torch

import alphastar_ref as O
import make_golden as G
from distar_b200.params import init_state_dict
from distar_b200.synth import tree_clone

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def close(a, b, name, rtol=1e-4):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    fin = b.abs() < 1e8
    assert torch.equal(fin, a.abs() < 1e8), name
    if fin.any():
        scale = max(b[fin].abs().max().item(), 1e-6)
        err = (a[fin] - b[fin]).abs().max().item()
        assert err <= rtol * scale, '%s: max err %.3e, scale %.3e' % (name, err, scale)


@pytest.fixture(scope='module')
def sd():
    return init_state_dict(seed=G.WEIGHT_SEED, baselines=G.BASELINES)


def test_fixture_inputs_reproduce(sd):
    g = torch.load(os.path.join(GOLD, 'infer.pt'))
    assert g['meta']['weights_checksum'] == G.checksum(sd), 'seeded weights differ on this machine'
    assert g['input_checksum'] == G.checksum(G.infer_case()), 'seeded inputs differ on this machine'


def test_oracle_sampling_forward_vs_golden(sd, su_action_mask):
    g = torch.load(os.path.join(GOLD, 'infer.pt'))
    torch.manual_seed(g['rng_seed'])
    with torch.no_grad():
        o = O.compute_logp_action(sd, **tree_clone(G.infer_case()), su_action_mask=su_action_mask)
    for k in O.HEADS:
        assert torch.equal(o['action_info'][k], g['action_info'][k]), k
        close(o['logit'][k], g['logit'][k], 'logit/' + k)
        close(o['action_logp'][k], g['action_logp'][k], 'logp/' + k)
    assert torch.equal(o['selected_units_num'], g['selected_units_num'])


def test_oracle_teacher_forward_vs_golden(sd):
    g = torch.load(os.path.join(GOLD, 'teacher.pt'))
    obs, act, num = G.teacher_case()
    with torch.no_grad():
        o = O.compute_teacher_logit(sd, **tree_clone(obs), selected_units_num=num, action_info=act)
    for k in O.HEADS:
        close(o['logit'][k], g['logit'][k], 'logit/' + k)


def test_oracle_rl_step_vs_golden(sd):
    g = torch.load(os.path.join(GOLD, 'rl_step.pt'))
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    out = O.rl_learner_forward(P, **tree_clone(G.rl_case()))
    info = O.rl_loss(out)
    info['total_loss'].backward()
    for k in O.HEADS:
        close(out['target_logit'][k], g['target_logit'][k], 'target_logit/' + k)
    for k, v in g['value'].items():
        close(out['value'][k], v, 'value/' + k)
    for k, v in g['loss'].items():
        assert abs(info[k].item() - v) <= 1e-4 * max(1.0, abs(v)), (k, info[k].item(), v)
    gmax = max(g['grad_norm'].values())
    for n, v in g['grad_norm'].items():
        assert abs(P[n].grad.norm().item() - v) <= 1e-3 * max(v, 1e-3 * gmax), n
    for n, v in g['grads'].items():
        close(P[n].grad, v, 'grad/' + n, rtol=1e-3)


def test_oracle_batch32_sampling_vs_golden(sd, su_action_mask):
    """BASELINE configs[1] at its size: every sampled index of the reference reproduced, logits through their compact form."""
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'infer32.pt'))
    assert g['input_checksum'] == G.checksum(G.infer32_case())
    torch.manual_seed(g['rng_seed'])
    with torch.no_grad():
        o = O.compute_logp_action(sd, **tree_clone(G.infer32_case()), su_action_mask=su_action_mask)
    for k in O.HEADS:
        assert torch.equal(o['action_info'][k], g['action_info'][k]), k
        assert_compact_close(o['logit'][k], g['logit'][k], 'logit/' + k, rtol=1e-4)
        close(o['action_logp'][k], g['action_logp'][k], 'logp/' + k)
    assert torch.equal(o['selected_units_num'], g['selected_units_num'])


def test_oracle_multi_chunk_rl_vs_golden(sd):
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'rl_chunks.pt'))
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    out = O.rl_learner_forward(P, **tree_clone(G.rl_chunk_case()))
    info = O.rl_loss(out)
    info['total_loss'].backward()
    for k in O.HEADS:
        assert_compact_close(out['target_logit'][k], g['target_logit'][k], 'target_logit/' + k, rtol=1e-4)
    for k, v in g['value'].items():
        close(out['value'][k], v, 'value/' + k)
    for k, v in g['loss'].items():
        assert abs(info[k].item() - v) <= 1e-4 * max(1.0, abs(v)), (k, info[k].item(), v)
    gmax = max(g['grad_norm'].values())
    for n, v in g['grad_norm'].items():
        assert abs(P[n].grad.norm().item() - v) <= 1e-3 * max(v, 1e-3 * gmax), n
        gen = torch.Generator().manual_seed(len(n) * 7919 + P[n].numel())
        pr = float((P[n].grad.reshape(-1) * torch.randn(P[n].numel(), generator=gen)).sum())
        assert abs(pr - g['grad_proj'][n]) <= 2e-3 * max(v, 1e-3 * gmax) * max(1.0, P[n].numel() ** 0.5 / 30), n
    for n, v in g['grads'].items():
        close(P[n].grad, v, 'grad/' + n, rtol=1e-3)


def test_oracle_value_feature_rl_vs_golden():
    """use_value_feature: True - the oracle's ValueEncoder restatement against what the reference produced."""
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'rl_value_feature.pt'))
    sd = init_state_dict(seed=G.VALUE_WEIGHT_SEED, baselines=G.VALUE_BASELINES, use_value_feature=True)
    assert G.checksum(sd) == g['meta']['weights_checksum'] and G.checksum(G.rl_value_case()) == g['input_checksum']
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    out = O.rl_learner_forward(P, **tree_clone(G.rl_value_case()))
    info = O.rl_loss(out)
    info['total_loss'].backward()
    for k in O.HEADS:
        assert_compact_close(out['target_logit'][k], g['target_logit'][k], 'target_logit/' + k, rtol=1e-4)
    for k, v in g['value'].items():
        close(out['value'][k], v, 'value/' + k)
    for k, v in g['loss'].items():
        assert abs(info[k].item() - v) <= 1e-4 * max(1.0, abs(v)), (k, info[k].item(), v)
    gmax = max(g['grad_norm'].values())
    for n, v in g['grad_norm'].items():
        assert abs(P[n].grad.norm().item() - v) <= 1e-3 * max(v, 1e-3 * gmax), n
    for n, v in g['grads'].items():
        close(P[n].grad, v, 'grad/' + n, rtol=1e-3)
    assert any(n.startswith('value_encoder.') for n in g['grads'])


def test_oracle_teacher_forward_default_map_size_vs_golden():
    """The reference's own default map size 160 x 152 (teacher-forced forward dumped from the real reference)."""
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'teacher_160x152.pt'))
    sx, sy = G.DEFAULT_XY
    sd = init_state_dict(seed=G.WEIGHT_SEED, spatial_x=sx, spatial_y=sy, baselines=G.BASELINES)
    obs, act, num = G.teacher_default_size_case()
    assert G.checksum(sd) == g['meta']['weights_checksum'] and G.checksum((obs, act, num)) == g['input_checksum']
    with torch.no_grad():
        o = O.compute_teacher_logit(sd, **tree_clone(obs), selected_units_num=num.clone(), action_info=tree_clone(act))
    for k in O.HEADS:
        assert_compact_close(o['logit'][k], g['logit'][k], 'logit/' + k, rtol=1e-4)
    assert o['logit']['target_location'].shape[-1] == sx * sy
    for (h, c), (gh, gc) in zip(o['hidden_state'], g['hidden_state']):
        close(h, gh, 'h')
        close(c, gc, 'c')
