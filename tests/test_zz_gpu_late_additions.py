"""-m gpu: parity tests of the two features that landed after this round's GPU allowance was used up - the ValueEncoder
(learner.use_value_feature) and the reference's default map size 160 x 152.  Both are pinned on the CPU (oracle vs the real
reference, oracle vs golden, product host logic vs oracle); their FIRST execution on a B200 is the driver's round-end run,
which is why they live in the last test file: nothing that was verified on the GPU during the round runs after them."""
import os

import pytest
import torch

import alphastar_ref as O
import make_golden as G
from distar_b200.model import Model
from distar_b200.params import init_state_dict
from distar_b200.rl_loss import ReinforcementLoss
from test_gpu_model import GOLD, _check_gradients, _oracle_gradients, close, to_dev

pytestmark = pytest.mark.gpu


def test_rl_step_with_value_feature_vs_golden():
    """learner.use_value_feature: True (the reference's self-play default): ValueEncoder (value_encoder.py:47-74) in front of
    the baselines, its spatial tower chunked (encoder_chunk=3 -> 3 chunks over 8 rows) and recomputed in backward."""
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'rl_value_feature.pt'))
    sd = init_state_dict(seed=G.VALUE_WEIGHT_SEED, baselines=G.VALUE_BASELINES, use_value_feature=True)
    cfg = {'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': list(G.VALUE_BASELINES)},
           'learner': {'use_value_feature': True}}
    m = Model(cfg, use_value_network=True, seed=0, encoder_chunk=3)
    m.load_state_dict(sd)
    m = m.cuda()
    m.zero_grad()
    batch = G.rl_value_case()
    out = m.rl_learner_forward(**to_dev(batch))
    info = ReinforcementLoss(None, 'MP0').compute_loss(out)
    info['total_loss'].backward()
    m.raise_on_bad_input()
    for k in O.HEADS:
        assert_compact_close(out['target_logit'][k], g['target_logit'][k], 'target_logit/' + k)
    for k, v in g['value'].items():
        close(out['value'][k], v, 'value/' + k)
    for k, v in g['loss'].items():
        got = info[k].item() if torch.is_tensor(info[k]) else info[k]
        assert abs(got - v) <= 2e-3 * max(1.0, abs(v)), (k, got, v)
    # 8 frames: ReLU / max-pool-decision sensitive like the 6-frame rl_step case, hence its (looser) floors
    _check_gradients(m, g, _oracle_gradients(sd, batch), cos_floor=0.995, frac_4nines=0.80)


def test_teacher_forward_default_map_size_vs_golden():
    """The reference's default map size 160 x 152: stem / pooling / 1x1 convolutions / up-sampling tail on our kernels, the 3x3
    convolutions at 76x80 / 38x40 / 19x20 on the fp32 library path (ops.conv_geometry_supported); forward + backward run."""
    from golden_util import assert_compact_close
    g = torch.load(os.path.join(GOLD, 'teacher_160x152.pt'))
    sx, sy = G.DEFAULT_XY
    sd = init_state_dict(seed=G.WEIGHT_SEED, spatial_x=sx, spatial_y=sy, baselines=G.BASELINES)
    m = Model({'model': {'spatial_x': sx, 'spatial_y': sy, 'enable_baselines': list(G.BASELINES)}}, use_value_network=True, seed=0)
    m.load_state_dict(sd)
    m = m.cuda()
    obs, act, num = G.teacher_default_size_case()
    o = m.compute_teacher_logit(**to_dev(obs), selected_units_num=num.cuda(), action_info=to_dev(act))
    for k in O.HEADS:
        assert_compact_close(o['logit'][k], g['logit'][k], 'logit/' + k)
    for (h, c), (gh, gc) in zip(o['hidden_state'], g['hidden_state']):
        close(h, gh, 'h')
        close(c, gc, 'c')
    m.zero_grad()
    sum(v[v > -1e8].square().mean() for v in o['logit'].values()).backward()
    gn = float(m.flat_grad.norm())
    assert gn == gn and gn > 0
