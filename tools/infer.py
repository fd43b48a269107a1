"""Dev helper: where does a batch-32 compute_logp_action call spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200.model import Model
from distar_b200.synth import synth_obs, tree_map
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0).cuda()
obs = tree_map(lambda t: t.to(dev), synth_obs(32, seed=7))
net = model._net()
def timeit(name, fn, reps=5):
    with torch.no_grad():
        out = fn(); torch.cuda.synchronize()
        t = time.time()
        for _ in range(reps): out = fn()
        torch.cuda.synchronize()
    print('%-28s %8.2f ms (wall)' % (name, (time.time() - t) / reps * 1e3)); return out
enc = timeit('encoder', lambda: net.encoder(obs['spatial_info'], obs['entity_info'], obs['scalar_info'], obs['entity_num']))
li, ctx, bf, ee, ms = enc
lo, st = timeit('core_lstm', lambda: net.lstm('core_lstm', li.unsqueeze(0), obs['hidden_state'], 3))
lo = lo.squeeze(0)
lg, at, emb = timeit('action_type_head', lambda: net.action_type_head(lo, ctx))
_, _, emb = timeit('delay_head', lambda: net.arg_head('policy.delay_head.', emb, 128, False))
_, _, emb = timeit('queued_head', lambda: net.arg_head('policy.queued_head.', emb, 2, True))
su_mask = torch.ones(32, dtype=torch.bool, device=dev)
r = timeit('selected_units_sample', lambda: net.selected_units_sample(emb, ee, obs['entity_num'], su_mask), reps=3)
print('  su steps run:', r[1].shape[1])
timeit('target_unit_head', lambda: net.target_unit_head(r[2], ee, obs['entity_num']))
timeit('location_head', lambda: net.location_head(r[2], ms))
timeit('compute_logp_action', lambda: model.compute_logp_action(**obs), reps=3)
