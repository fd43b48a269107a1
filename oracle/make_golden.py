"""Dump golden fixtures from the UNMODIFIED reference (run in the authoring container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt

Weights come from distar_b200.params.init_state_dict (seeded, reproducible anywhere) loaded into the reference
``Model``; inputs from distar_b200.synth (seeded).  The fixtures hold only reference OUTPUTS (plus input
checksums), so the GPU box — which has no /root/reference — can check the oracle and the CUDA path against
what the real reference produced.
"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import ref_import  # noqa: E402
from distar_b200.params import init_state_dict  # noqa: E402
from distar_b200.synth import synth_obs, synth_rl_batch, synth_actions, tree_clone, tree_map  # noqa: E402
from golden_util import compact_logits  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
BASELINES = ('winloss', 'build_order')
WEIGHT_SEED = 3


def checksum(tree) -> str:
    h = hashlib.sha256()

    def visit(t):
        h.update(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
        return t
    tree_map(visit, tree)
    return h.hexdigest()[:16]


def infer_case():
    return synth_obs(3, seed=11, entity_num=torch.tensor([512, 77, 300]))


def teacher_case():
    en = torch.tensor([512, 40, 333, 200])
    obs = synth_obs(4, seed=12, entity_num=en)
    g = torch.Generator().manual_seed(1)
    act, num = synth_actions(4, en, g, max_su=9)
    num[1] = 0
    return obs, act, num


def rl_case():
    batch = synth_rl_batch(2, 3, seed=21, entity_num='random', max_su=6)
    batch['reward']['winloss'][-1, 0] = 1.0
    return batch


def infer32_case():
    """BASELINE configs[1]: batch-32 inference (compute_logp_action), ragged entity counts."""
    return synth_obs(32, seed=51, entity_num='random')


def rl_chunk_case():
    """An RL batch whose (T+1)*B = 36 observation rows span three encoder chunks at encoder_chunk=16."""
    batch = synth_rl_batch(4, 8, seed=61, entity_num='random', max_su=6)
    batch['reward']['winloss'][-1, 1] = -1.0
    return batch


DEFAULT_XY = (160, 152)          # the reference's own default map size (actor_critic_default_config.yaml: spatial_x / spatial_y)


def teacher_default_size_case():
    """Teacher-forced forward at 160 x 152 (maps [152, 160]; 76x80 / 38x40 / 19x20 inside the spatial tower)."""
    en = torch.tensor([512, 40, 333])
    obs = synth_obs(3, seed=81, entity_num=en, hw=(DEFAULT_XY[1], DEFAULT_XY[0]))
    g = torch.Generator().manual_seed(3)
    act, num = synth_actions(3, en, g, max_su=7, hw=(DEFAULT_XY[1], DEFAULT_XY[0]))
    return obs, act, num


def default_size_golden():
    model, cfg, mods = ref_import.load_reference(spatial=DEFAULT_XY, enable_baselines=BASELINES)
    sd = init_state_dict(seed=WEIGHT_SEED, spatial_x=DEFAULT_XY[0], spatial_y=DEFAULT_XY[1], baselines=BASELINES)
    model.load_state_dict(sd, strict=True)
    model.eval()
    meta = {'weight_seed': WEIGHT_SEED, 'baselines': list(BASELINES), 'spatial_xy': DEFAULT_XY, 'weights_checksum': checksum(sd),
            'torch': str(torch.__version__)}
    obs, act, num = teacher_default_size_case()
    with torch.no_grad():
        r = model.compute_teacher_logit(**tree_clone(obs), selected_units_num=num.clone(), action_info=tree_clone(act))
    torch.save({'meta': meta, 'input_checksum': checksum((obs, act, num)),
                'logit': {k: compact_logits(v) for k, v in r['logit'].items()}, 'hidden_state': r['hidden_state']},
               os.path.join(OUT, 'teacher_160x152.pt'))


def rl_value_case():
    """learner.use_value_feature: True (bin/rl_user_config.yaml self-play default): 8 rows with the ValueEncoder inputs."""
    return synth_rl_batch(2, 3, seed=71, entity_num='random', max_su=6, value_feature=True)


VALUE_BASELINES = ('winloss', 'battle')
VALUE_WEIGHT_SEED = 4

FULL_GRADS = ['value_encoder.encode_modules.unit_type.weight', 'value_encoder.project.0.weight',
              'value_encoder.spatial_fc.0.bias', 'policy.action_type_head.action_fc.layer2.0.bias', 'core_lstm.layers.2.cell.layernorm_c.weight',
              'encoder.scatter_project.0.weight', 'value_networks.winloss.value_fc.0.weight',
              'policy.selected_units_head.end_embedding', 'encoder.spatial_encoder.project.0.weight']


def dump_rl(model, loss_fn, batch, meta, path, compact):
    model.zero_grad()
    out = model.rl_learner_forward(**tree_clone(batch))
    info = loss_fn.compute_loss(out)
    info['total_loss'].backward()
    scalars = {k: (v.item() if torch.is_tensor(v) else float(v)) for k, v in info.items()}
    grad_norm = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.requires_grad}
    # a second scalar per tensor: its projection on a seeded random direction (pins the direction, not only the length)
    proj = {}
    for n, p in model.named_parameters():
        if p.requires_grad:
            g = torch.Generator().manual_seed(len(n) * 7919 + p.numel())
            proj[n] = float((p.grad.reshape(-1) * torch.randn(p.numel(), generator=g)).sum())
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if n in FULL_GRADS}
    logits = {k: (compact_logits(v) if compact else v.detach()) for k, v in out['target_logit'].items()}
    torch.save({'meta': meta, 'input_checksum': checksum(batch), 'target_logit': logits, 'compact': compact,
                'value': {k: v.detach() for k, v in out['value'].items()}, 'loss': scalars,
                'grad_norm': grad_norm, 'grad_proj': proj, 'grads': grads}, path)


def value_feature_golden():
    model, cfg, mods = ref_import.load_reference(spatial=128, enable_baselines=VALUE_BASELINES, use_value_feature=True)
    sd = init_state_dict(seed=VALUE_WEIGHT_SEED, baselines=VALUE_BASELINES, use_value_feature=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    meta = {'weight_seed': VALUE_WEIGHT_SEED, 'baselines': list(VALUE_BASELINES), 'use_value_feature': True,
            'weights_checksum': checksum(sd), 'torch': str(torch.__version__)}
    dump_rl(model, mods['ReinforcementLoss'](cfg.learner, 'MP0'), rl_value_case(), meta, os.path.join(OUT, 'rl_value_feature.pt'),
            compact=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    if '--only-value-feature' in sys.argv:
        value_feature_golden()
        print('rl_value_feature.pt', os.path.getsize(os.path.join(OUT, 'rl_value_feature.pt')))
        return
    if '--only-default-size' in sys.argv:
        default_size_golden()
        print('teacher_160x152.pt', os.path.getsize(os.path.join(OUT, 'teacher_160x152.pt')))
        return
    model, cfg, mods = ref_import.load_reference(spatial=128, enable_baselines=BASELINES)
    sd = init_state_dict(seed=WEIGHT_SEED, baselines=BASELINES)
    model.load_state_dict(sd, strict=True)
    model.eval()
    meta = {'weight_seed': WEIGHT_SEED, 'baselines': list(BASELINES), 'weights_checksum': checksum(sd),
            'torch': str(torch.__version__)}
    # ---- config 1/2: sampling forward
    obs = infer_case()
    torch.manual_seed(5)
    with torch.no_grad():
        r = model.compute_logp_action(**tree_clone(obs))
    torch.save({'meta': meta, 'input_checksum': checksum(obs), 'rng_seed': 5,
                'action_info': r['action_info'], 'action_logp': r['action_logp'],
                'selected_units_num': r['selected_units_num'], 'logit': r['logit'],
                'hidden_state': r['hidden_state']}, os.path.join(OUT, 'infer.pt'))
    # ---- teacher-forced forward
    obs, act, num = teacher_case()
    with torch.no_grad():
        r = model.compute_teacher_logit(**tree_clone(obs), selected_units_num=num.clone(), action_info=tree_clone(act))
    torch.save({'meta': meta, 'input_checksum': checksum((obs, act, num)), 'logit': r['logit'],
                'hidden_state': r['hidden_state']}, os.path.join(OUT, 'teacher.pt'))
    # ---- config 4: RL step (6 frames), and a 36-row batch that spans several encoder chunks
    loss_fn = mods['ReinforcementLoss'](cfg.learner, 'MP0')
    dump_rl(model, loss_fn, rl_case(), meta, os.path.join(OUT, 'rl_step.pt'), compact=False)
    dump_rl(model, loss_fn, rl_chunk_case(), meta, os.path.join(OUT, 'rl_chunks.pt'), compact=True)
    # ---- config 2: batch-32 sampling forward (compact logits)
    obs = infer32_case()
    torch.manual_seed(7)
    with torch.no_grad():
        r = model.compute_logp_action(**tree_clone(obs))
    torch.save({'meta': meta, 'input_checksum': checksum(obs), 'rng_seed': 7,
                'action_info': r['action_info'], 'action_logp': r['action_logp'],
                'selected_units_num': r['selected_units_num'],
                'logit': {k: compact_logits(v) for k, v in r['logit'].items()},
                'hidden_state': r['hidden_state']}, os.path.join(OUT, 'infer32.pt'))
    value_feature_golden()
    default_size_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
