"""Pins the oracle restatement (oracle/alphastar_ref.py) against the UNMODIFIED reference executed here.

Runs only where /root/reference exists (the authoring container); the GPU box relies on tests/golden/.
"""
import pytest
import torch

import ref_import
import alphastar_ref as O
from distar_b200.params import init_state_dict
from distar_b200.synth import synth_obs, synth_rl_batch, synth_actions, tree_clone

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_import.reference_available(), reason='reference tree not mounted')]


@pytest.fixture(scope='module')
def ref():
    model, cfg, mods = ref_import.load_reference(spatial=128, enable_baselines=('winloss', 'build_order'))
    sd = init_state_dict(seed=3, baselines=('winloss', 'build_order'))
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    model.eval()
    return model, cfg, mods, sd


def _close(a, b, rtol=1e-4, atol=None, name=''):
    a, b = a.detach().float(), b.detach().float()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    fin = b.abs() < 1e8
    assert torch.equal(fin, a.abs() < 1e8), name
    scale = b[fin].abs().max().item() if fin.any() else 1.0
    atol = atol if atol is not None else 1e-5 * max(scale, 1.0)
    assert torch.allclose(a[fin], b[fin], rtol=rtol, atol=atol), \
        '%s: max abs err %.3e (scale %.3e)' % (name, (a[fin] - b[fin]).abs().max().item(), scale)
    assert torch.equal(a[~fin], b[~fin]), name


def test_sampling_forward_matches_reference(ref, su_action_mask):
    model, cfg, mods, sd = ref
    obs = synth_obs(3, seed=11, entity_num=torch.tensor([512, 77, 300]))
    torch.manual_seed(5)
    with torch.no_grad():
        r = model.compute_logp_action(**tree_clone(obs))
    torch.manual_seed(5)
    with torch.no_grad():
        o = O.compute_logp_action(sd, **tree_clone(obs), su_action_mask=su_action_mask)
    for k in O.HEADS:
        assert torch.equal(r['action_info'][k], o['action_info'][k]), k
        _close(o['logit'][k], r['logit'][k], name='logit/' + k)
        _close(o['action_logp'][k], r['action_logp'][k], name='logp/' + k)
    assert torch.equal(r['selected_units_num'], o['selected_units_num'])
    for (rh, rc), (oh, oc) in zip(r['hidden_state'], o['hidden_state']):
        _close(oh, rh, name='h')
        _close(oc, rc, name='c')


def test_teacher_forward_matches_reference(ref):
    model, cfg, mods, sd = ref
    en = torch.tensor([512, 40, 333, 200])
    obs = synth_obs(4, seed=12, entity_num=en)
    g = torch.Generator().manual_seed(1)
    act, num = synth_actions(4, en, g, max_su=9)
    num[1] = 0
    with torch.no_grad():
        r = model.compute_teacher_logit(**tree_clone(obs), selected_units_num=num.clone(), action_info=tree_clone(act))
        o = O.compute_teacher_logit(sd, **tree_clone(obs), selected_units_num=num.clone(), action_info=tree_clone(act))
    for k in O.HEADS:
        _close(o['logit'][k], r['logit'][k], name='logit/' + k)


def test_rl_step_matches_reference(ref):
    model, cfg, mods, sd = ref
    batch = synth_rl_batch(2, 3, seed=21, entity_num='random', max_su=6)
    batch['reward']['winloss'][-1, 0] = 1.0
    loss_fn = mods['ReinforcementLoss'](cfg.learner, 'MP0')
    model.zero_grad()
    r_out = model.rl_learner_forward(**tree_clone(batch))
    r_info = loss_fn.compute_loss(r_out)
    r_info['total_loss'].backward()
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    o_out = O.rl_learner_forward(P, **tree_clone(batch))
    o_info = O.rl_loss(o_out)
    o_info['total_loss'].backward()
    for k in O.HEADS:
        _close(o_out['target_logit'][k], r_out['target_logit'][k], name='target_logit/' + k)
    for k in r_out['value']:
        _close(o_out['value'][k], r_out['value'][k], name='value/' + k)
    for k, v in r_info.items():
        rv = v.item() if torch.is_tensor(v) else v
        assert k in o_info, k
        assert abs(o_info[k].item() - rv) <= 1e-4 * max(1.0, abs(rv)), (k, o_info[k].item(), rv)
    n_checked = 0
    gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.requires_grad)
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        _close(P[name].grad, p.grad, rtol=1e-3, atol=1e-4 * max(p.grad.abs().max().item(), 1e-4 * gmax), name='grad/' + name)
        n_checked += 1
    assert n_checked > 300


def test_sl_forward_matches_reference(ref):
    model, cfg, mods, sd = ref
    B, T = 2, 2
    en = torch.tensor([512, 100, 256, 64])
    obs = synth_obs(B * T, seed=31, entity_num=en, hidden=False)
    g = torch.Generator().manual_seed(2)
    act, num = synth_actions(B * T, en, g, max_su=5)
    hidden = [(torch.randn(B, 384, generator=g), torch.randn(B, 384, generator=g)) for _ in range(3)]
    with torch.no_grad():
        rl, ra, rs = model.sl_train(**tree_clone(obs), selected_units_num=num.clone(), traj_lens=[T] * B,
                                    hidden_state=tree_clone(hidden), action_info=tree_clone(act))
        ol, oa, os_ = O.sl_train(sd, **tree_clone(obs), selected_units_num=num.clone(), traj_lens=[T] * B,
                                 hidden_state=tree_clone(hidden), action_info=tree_clone(act))
    for k in O.HEADS:
        _close(ol[k], rl[k], name='sl_logit/' + k)
    amask = {k: (torch.rand(B * T, generator=g) < 0.7).float() for k in O.HEADS}
    sl = mods['SupervisedLoss']({'learner': {'su_mask': False}})
    r_loss = sl.compute_loss(rl, act, amask, num, en, ra)
    assert oa['selected_units'] is None and ra['selected_units'] is None
    o_loss = O.sl_loss(ol, act, amask, num)
    for k, v in o_loss.items():
        assert abs(v.item() - float(r_loss[k])) <= 1e-4 * max(1.0, abs(float(r_loss[k]))), k
