"""-m gpu: the CUDA-graph inference server and the arena weight publication against eager calls."""
import pytest
import torch

import alphastar_ref as O
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.params import init_state_dict
from distar_b200.serving import InferenceServer, WeightPublisher, WeightSubscriber
from distar_b200.synth import synth_obs, synth_rl_batch, tree_clone, tree_map

pytestmark = pytest.mark.gpu
DEV = 'cuda'
CFG = {'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}


def to_dev(tree):
    return tree_map(lambda t: t.to(DEV), tree)


def close(a, b, name, rtol=1e-4):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    fin = b.abs() < 1e8
    assert torch.equal(fin, a.abs() < 1e8), name
    if fin.any():
        assert (a[fin] - b[fin]).abs().max().item() <= rtol * max(b[fin].abs().max().item(), 1e-6), name


def _actor(seed):
    m = Model(CFG, use_value_network=False, seed=0)
    m.load_state_dict({k: v for k, v in init_state_dict(seed=seed).items() if not k.startswith('value_networks')}, strict=False)
    return m.cuda()


def test_graph_replay_matches_eager_calls():
    actor, teacher = _actor(3), _actor(4)
    obs = synth_obs(4, seed=9, entity_num=torch.tensor([512, 77, 300, 5]))
    server = InferenceServer(actor, obs, teacher=teacher, su_steps=16)
    for seed in (9, 10):                                   # the second request reuses the captured graph with new inputs
        obs = synth_obs(4, seed=seed, entity_num=torch.tensor([512, 77, 300, 5]))
        out = server.infer(tree_map(lambda t: t.pin_memory(), obs), teacher_hidden=obs['hidden_state'])
        steps = out['action_info']['selected_units'].shape[1]
        assert steps == max(1, int(out['selected_units_num'].max())) and out['logit']['selected_units'].shape[1] == steps
        with torch.no_grad():
            eager = actor.compute_logp_action(**to_dev(obs))
            # everything up to the first sampling step is deterministic: identical logits; sampled ids are valid
            close(out['logit']['action_type'], eager['logit']['action_type'], 'action_type logits')
            assert bool((out['action_info']['action_type'] < 327).all()) and bool((out['action_info']['target_unit'] < 512).all())
            for (h, c), (eh, ec) in zip(out['hidden_state'], eager['hidden_state']):
                close(h, eh, 'h')
                close(c, ec, 'c')
            # the teacher's forced pass on the SAMPLED action must equal an eager teacher call on the same action
            t = teacher.compute_teacher_logit(**to_dev(obs), selected_units_num=out['selected_units_num'],
                                              action_info=out['action_info'])
        for k in O.HEADS:
            want = t['logit'][k]
            if k == 'selected_units':
                want = want[:, :steps]
                got = out['teacher_logit'][k][:, :want.shape[1]]
                close(got, want, 'teacher/' + k)
            else:
                close(out['teacher_logit'][k], want, 'teacher/' + k)
        # log-probs are those of the returned logits
        for k in ('action_type', 'delay', 'queued', 'target_unit', 'target_location'):
            lp = torch.log_softmax(out['logit'][k], -1).gather(-1, out['action_info'][k].unsqueeze(-1)).squeeze(-1)
            close(out['action_logp'][k], lp, 'logp/' + k)
    assert server.replays == 2


def test_weight_publication_reaches_the_captured_graph():
    """learner arena -> pinned snapshot -> actor arena (different layout: no value networks) -> in-place refresh of the derived
    weight forms: the SAME captured graph must now compute with the new weights."""
    learner_model = Model(CFG, use_value_network=True, seed=0)
    learner_model.load_state_dict(init_state_dict(seed=5))
    learner_model = learner_model.cuda()
    actor = _actor(3)
    obs = synth_obs(2, seed=21, entity_num=torch.tensor([40, 512]))
    server = InferenceServer(actor, obs, su_steps=8)
    before = server.infer(obs)['logit']['action_type'].clone()
    pub, sub = WeightPublisher(learner_model), WeightSubscriber(actor, server)
    assert 'value_networks.winloss.project.0.weight' not in pub.layout
    learner = RLLearner(learner_model, 'MP0', lr=1e-3)
    learner._train(to_dev(synth_rl_batch(1, 2, seed=3, max_su=4)))
    v = pub.publish()
    sub.update(pub)
    assert sub.version == v == 1
    for n, p in actor.named_parameters():
        if p.requires_grad:
            assert torch.equal(p, dict(learner_model.named_parameters())[n]), n
    after = server.infer(obs)['logit']['action_type']
    fresh = _actor(3)
    fresh.load_state_dict({k: v for k, v in learner_model.state_dict().items() if not k.startswith('value_networks')}, strict=False)
    with torch.no_grad():
        want = fresh.compute_logp_action(**to_dev(obs))['logit']['action_type']
    close(after, want, 'after publication')
    assert (after - before).abs().max().item() > 1e-4          # and they did change
