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


def test_dapo_loss_matches_reference(ref):
    """rl_loss.py:164-172 / as_rl_utils.py:105-127: the 'MP' players' KL towards the successive model's logits."""
    model, cfg, mods, sd = ref
    batch = synth_rl_batch(2, 3, seed=23, entity_num='random', max_su=5)
    g = torch.Generator().manual_seed(9)
    succ = {k: (v + 0.5 * torch.randn(v.shape, generator=g)).masked_fill(v < -1e8, -1e9) for k, v in batch['teacher_logit'].items()}
    batch['step'][0, 0] = 100.0                                         # at least one frame inside dapo_steps
    import copy
    lcfg = copy.deepcopy(cfg.learner)
    lcfg.use_dapo = True
    loss_fn = mods['ReinforcementLoss'](lcfg, 'MP0')
    with torch.no_grad():
        r_out = model.rl_learner_forward(**tree_clone(batch))
        r_out['successive_logit'] = tree_clone(succ)
        r_info = loss_fn.compute_loss(r_out)
        o_out = O.rl_learner_forward(sd, **tree_clone(batch))
        o_out['successive_logit'] = tree_clone(succ)
        o_info = O.rl_loss(o_out, use_dapo=True, dapo_w=float(lcfg.loss_weights.dapo), dapo_steps=int(lcfg.dapo.dapo_steps))
    for k, v in r_info.items():
        rv = v.item() if torch.is_tensor(v) else v
        assert abs(o_info[k].item() - rv) <= 1e-4 * max(1.0, abs(rv)), (k, o_info[k].item(), rv)
    assert abs(r_info['battle/total']) > 0
    # non-'MP' players never use it (rl_loss.py:22-24)
    assert mods['ReinforcementLoss'](lcfg, 'EP0').use_dapo is False


@pytest.mark.parametrize('su_mask,label_smooth', [(True, False), (False, True), (True, True)])
def test_sl_loss_options_match_reference(ref, su_mask, label_smooth):
    """sl_loss.py:54-57 (label smoothing), :177-192 (su_mask), :206-232 (IoU) on seeded logits."""
    model, cfg, mods, sd = ref
    g = torch.Generator().manual_seed(17)
    b, s, E = 6, 5, 512
    en = torch.tensor([512, 40, 333, 200, 64, 7])
    act, num = synth_actions(b, en, g, max_su=5)
    valid = torch.arange(E + 1).unsqueeze(0) < (en + 1).unsqueeze(1)
    logits = {'action_type': torch.randn(b, 327, generator=g), 'delay': torch.randn(b, 128, generator=g),
              'queued': torch.randn(b, 2, generator=g),
              'selected_units': torch.randn(b, s, E + 1, generator=g).masked_fill(~valid.unsqueeze(1), -1e9),
              'target_unit': torch.randn(b, E, generator=g).masked_fill(~valid[:, :E], -1e9),
              'target_location': torch.randn(b, 128 * 128, generator=g)}
    amask = {k: (torch.rand(b, generator=g) < 0.7).float() for k in O.HEADS}
    # a plausible sampled selection: the labels with one unit swapped, end token kept
    preds = act['selected_units'][:, :s].clone()
    preds[:, 0] = (preds[:, 0] + 1) % en.clamp(min=2)
    sl = mods['SupervisedLoss']({'learner': {'su_mask': su_mask, 'label_smooth': label_smooth}})
    r = sl.compute_loss(tree_clone(logits), tree_clone(act), tree_clone(amask), num.clone(), en.clone(),
                        {'selected_units': preds.clone()})
    o = O.sl_loss(tree_clone(logits), tree_clone(act), tree_clone(amask), num.clone(), en.clone(), preds.clone(),
                  su_mask=su_mask, label_smooth=label_smooth)
    assert set(o.keys()) == set(r.keys())
    for k, v in r.items():
        assert abs(float(o[k]) - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (k, float(o[k]), float(v))
    assert mods['SupervisedLoss']({'learner': {}}).su_mask is True          # default_supervised_loss.yaml:10


def test_momentum_norm_clip_never_scales(ref):
    """grad_clip.py:74-107 as written: the norm momenta are appended instead of stored at their index, so no gradient is
    ever scaled and apply() returns the global 2-norm.  ops.FlatAdam(clip_type='momentum_norm') relies on exactly this."""
    ref_import.install_shims()
    from distar.ctools.torch_utils.grad_clip import build_grad_clip
    clip = build_grad_clip({'type': 'momentum_norm', 'threshold': 1.4})
    g = torch.Generator().manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(7, 5, generator=g)), torch.nn.Parameter(torch.randn(11, generator=g))]
    for it in range(4):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g) * (10.0 ** it)      # growing norms would trigger a working clip
        before = [p.grad.clone() for p in params]
        total = clip.apply(params)
        want = torch.sqrt(sum((b ** 2).sum() for b in before)).item()
        assert abs(total - want) <= 1e-5 * want
        for p, b in zip(params, before):
            assert torch.equal(p.grad, b)


def test_adam_state_interchanges_with_reference_optimizer(ref):
    """rl_learner.py:73-79 builds torch.optim.Adam over model.parameters(); checkpoints carry its state_dict()
    (checkpoint_helper.py:124-131).  FlatAdam must read it, continue identically, and write it back in the same layout."""
    from distar_b200 import ops
    from distar_b200.model import Model
    model, cfg, mods, sd = ref
    ops.enable_host_logic_testing(True)
    try:
        mine = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss', 'build_order']}},
                     use_value_network=True, seed=0)
        mine.load_state_dict(sd)
        ref_model = mods['Model'](cfg, use_value_network=True)
        ref_model.load_state_dict(sd, strict=True)
        opt_ref = torch.optim.Adam(ref_model.parameters(), lr=1e-3, betas=(0.0, 0.99), eps=1e-5)
        gen = torch.Generator().manual_seed(0)
        grads = {n: torch.randn(p.shape, generator=gen) for n, p in ref_model.named_parameters() if p.requires_grad}

        def set_grads(m):
            for n, p in m.named_parameters():
                if p.requires_grad:
                    p.grad.copy_(grads[n]) if p.grad is not None else setattr(p, 'grad', grads[n].clone())
        set_grads(ref_model)
        opt_ref.step()
        # hand the reference's optimizer state + weights to FlatAdam and take the SECOND step on both sides
        opt = ops.FlatAdam(mine.flat_param, mine.flat_grad, lr=1e-3, betas=(0.0, 0.99), eps=1e-5, max_norm=None, clip_type="none",
                           layout=mine.optimizer_layout(), owner=mine)
        mine.load_state_dict(ref_model.state_dict())
        opt.load_state_dict(opt_ref.state_dict())
        assert opt.t == 1
        set_grads(ref_model)
        set_grads(mine)
        opt_ref.step()
        opt.step()
        for (n, a), (_, b) in zip(mine.named_parameters(), ref_model.named_parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), n
        back, want = opt.state_dict(), opt_ref.state_dict()
        assert back['param_groups'][0]['params'] == want['param_groups'][0]['params']
        assert set(back['state'].keys()) == set(want['state'].keys())
        for i, st in want['state'].items():
            assert float(back['state'][i]['step']) == float(st['step']) == 2
            assert torch.allclose(back['state'][i]['exp_avg_sq'], st['exp_avg_sq'], rtol=1e-5, atol=1e-9), i
            assert torch.allclose(back['state'][i]['exp_avg'], st['exp_avg'], rtol=1e-5, atol=1e-9), i
    finally:
        ops.enable_host_logic_testing(False)


def test_rl_step_with_value_feature_matches_reference():
    """learner.use_value_feature: True (the bin/rl_user_config.yaml default for self-play): ValueEncoder
    (obs_encoder/value_encoder.py) feeds every baseline (model.py:141-144, value.py:20-23)."""
    model, cfg, mods = ref_import.load_reference(spatial=128, enable_baselines=('winloss', 'battle'), use_value_feature=True)
    sd = init_state_dict(seed=4, baselines=('winloss', 'battle'), use_value_feature=True)
    model.load_state_dict(sd, strict=True)
    assert [k for k, _ in model.named_parameters()] == [k for k in sd if dict(model.named_parameters()).get(k) is not None]
    batch = synth_rl_batch(2, 2, seed=23, entity_num='random', max_su=5, value_feature=True)
    loss_fn = mods['ReinforcementLoss'](cfg.learner, 'MP0')
    model.zero_grad()
    r_out = model.rl_learner_forward(**tree_clone(batch))
    r_info = loss_fn.compute_loss(r_out)
    r_info['total_loss'].backward()
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    o_out = O.rl_learner_forward(P, **tree_clone(batch))
    o_info = O.rl_loss(o_out)
    o_info['total_loss'].backward()
    for k in r_out['value']:
        _close(o_out['value'][k], r_out['value'][k], name='value/' + k)
    for k, v in r_info.items():
        rv = v.item() if torch.is_tensor(v) else v
        assert abs(o_info[k].item() - rv) <= 1e-4 * max(1.0, abs(rv)), (k, o_info[k].item(), rv)
    gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.requires_grad)
    n_ve = 0
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        _close(P[name].grad, p.grad, rtol=1e-3, atol=1e-4 * max(p.grad.abs().max().item(), 1e-4 * gmax), name='grad/' + name)
        n_ve += name.startswith('value_encoder.')
    assert n_ve > 60


def test_default_map_size_matches_reference(su_action_mask):
    """160 x 152 (the reference's default spatial_x / spatial_y): sampling forward and one RL step, oracle vs the real reference."""
    sx, sy = 160, 152
    model, cfg, mods = ref_import.load_reference(spatial=(sx, sy), enable_baselines=('winloss',))
    sd = init_state_dict(seed=5, spatial_x=sx, spatial_y=sy, baselines=('winloss',))
    model.load_state_dict(sd, strict=True)
    model.eval()
    obs = synth_obs(2, seed=91, entity_num=torch.tensor([300, 512]), hw=(sy, sx))
    torch.manual_seed(9)
    with torch.no_grad():
        r = model.compute_logp_action(**tree_clone(obs))
    torch.manual_seed(9)
    with torch.no_grad():
        o = O.compute_logp_action(sd, **tree_clone(obs), su_action_mask=su_action_mask)
    for k in O.HEADS:
        assert torch.equal(r['action_info'][k], o['action_info'][k]), k
        _close(o['logit'][k], r['logit'][k], name='logit/' + k)
    assert o['logit']['target_location'].shape[-1] == sx * sy
    # one RL step: the batch generator only knows 128 x 128, so swap in 160 x 152 observations and location labels / teachers
    batch = synth_rl_batch(2, 2, seed=25, entity_num='random', max_su=5)
    big = synth_obs(6, seed=25, entity_num=batch['entity_num'], hw=(sy, sx))
    batch['spatial_info'], batch['scalar_info']['bo_location'] = big['spatial_info'], big['scalar_info']['bo_location']
    batch['entity_info']['x'], batch['entity_info']['y'] = big['entity_info']['x'], big['entity_info']['y']
    g = torch.Generator().manual_seed(3)
    batch['action_info']['target_location'] = torch.randint(0, sx * sy, (2, 2), generator=g)
    batch['teacher_logit']['target_location'] = torch.randn(2, 2, sx * sy, generator=g)
    loss_fn = mods['ReinforcementLoss'](cfg.learner, 'MP0')
    model.zero_grad()
    r_info = loss_fn.compute_loss(model.rl_learner_forward(**tree_clone(batch)))
    r_info['total_loss'].backward()
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    o_info = O.rl_loss(O.rl_learner_forward(P, **tree_clone(batch)))
    o_info['total_loss'].backward()
    for k, v in r_info.items():
        rv = v.item() if torch.is_tensor(v) else v
        assert abs(o_info[k].item() - rv) <= 1e-4 * max(1.0, abs(rv)), (k, o_info[k].item(), rv)
    gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.requires_grad)
    for name, p in model.named_parameters():
        if p.requires_grad:
            _close(P[name].grad, p.grad, rtol=1e-3, atol=1e-4 * max(p.grad.abs().max().item(), 1e-4 * gmax), name='grad/' + name)


def test_only_update_baseline_with_value_feature_product_vs_reference(monkeypatch):
    """model.only_update_baseline: True (model.py:138-140): the critic reads DETACHED lstm output / baseline feature, so the value
    losses train only the value networks and the ValueEncoder.  Product host logic (kernel stand-ins, exact operand split)
    against the real reference: values and every gradient."""
    from distar_b200 import ops
    from distar_b200.model import Model
    from distar_b200.rl_loss import ReinforcementLoss
    ref_model, cfg, mods = ref_import.load_reference(spatial=128, enable_baselines=('winloss',), use_value_feature=True)
    ref_model.only_update_baseline = True
    sd = init_state_dict(seed=6, baselines=('winloss',), use_value_feature=True)
    ref_model.load_state_dict(sd, strict=True)
    batch = synth_rl_batch(2, 2, seed=27, entity_num='random', max_su=5, value_feature=True)
    ref_model.zero_grad()
    r_out = ref_model.rl_learner_forward(**tree_clone(batch))
    r_info = mods['ReinforcementLoss'](cfg.learner, 'MP0').compute_loss(r_out)
    r_info['total_loss'].backward()
    ops.enable_host_logic_testing(True)
    monkeypatch.setattr(ops, 'split_bf16', lambda x: (x.contiguous(), torch.zeros_like(x)))
    try:
        mine = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss'], 'only_update_baseline': True},
                      'learner': {'use_value_feature': True}}, use_value_network=True, seed=0)
        mine.load_state_dict(sd)
        mine.zero_grad()
        out = mine.rl_learner_forward(**tree_clone(batch))
        info = ReinforcementLoss(None, 'MP0').compute_loss(out)
        info['total_loss'].backward()
    finally:
        ops.enable_host_logic_testing(False)
    _close(out['value']['winloss'], r_out['value']['winloss'], rtol=1e-3, name='value')
    ref_grads = dict(ref_model.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in ref_model.parameters() if p.requires_grad and p.grad is not None)
    for n, p in mine.named_parameters():
        if not p.requires_grad:
            continue
        rg = ref_grads[n].grad
        rg = torch.zeros_like(p) if rg is None else rg
        assert (p.grad - rg).abs().max().item() <= 2e-3 * max(rg.abs().max().item(), 1e-3 * gmax), n
