"""Dev helper: aggregated per-kernel device time of one learner step (torch.profiler, not a number of record)."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
B, T = int(sys.argv[1]), int(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 264
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0, encoder_chunk=chunk, checkpoint_encoder=True, keep_chunks=int(os.environ.get('KEEP', 13))).cuda()
learner = RLLearner(model)
data = tree_map(lambda t: t.to(dev), synth_rl_batch(B, T, seed=0))
learner._train(data); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    learner._train(data); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name[:110]][0] += 1; agg[ev.name[:110]][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print('total device us', tot, 'kernels', sum(v[0] for v in agg.values()))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:70]:
    print('%10.0f us %5.1f%% n=%6d %s' % (t, 100 * t / tot, n, k))
