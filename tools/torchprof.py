"""Dev helper: aggregated per-kernel device time of one learner step (torch.profiler, not a number of record)."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
B, T = int(sys.argv[1]), int(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 264
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0, encoder_chunk=chunk, checkpoint_encoder=True, keep_chunks=int(os.environ.get('KEEP', 16))).cuda()
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
# ---- where does the GPU idle?  gaps between consecutive kernels, attributed to the kernel that ends the gap
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
gap_by = collections.defaultdict(lambda: [0, 0.0])
end = evs[0].time_range.end
t_first = evs[0].time_range.start
timeline = []
for e in evs[1:]:
    g = e.time_range.start - end
    if g > 2:
        gap_by[e.name[:90]][0] += 1; gap_by[e.name[:90]][1] += g
        timeline.append((e.time_range.start - t_first, g, e.name[:60]))
    end = max(end, e.time_range.end)
print('span us', end - t_first, 'idle us', sum(v[1] for v in gap_by.values()))
for k, (n, t) in sorted(gap_by.items(), key=lambda x: -x[1][1])[:25]:
    print('%9.0f us idle before n=%5d %s' % (t, n, k))
# idle per 20 ms window of the step
win = collections.defaultdict(float)
for t, g, _ in timeline:
    win[int(t // 20000)] += g
print('idle ms per 20 ms window:', ' '.join('%d:%.1f' % (w, win[w] / 1e3) for w in sorted(win)))
