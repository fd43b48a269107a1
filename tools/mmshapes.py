"""Dev helper: which library matmuls (aten::mm / addmm / bmm / matmul) are left in a learner step, by input shape."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0,
              encoder_chunk=264, checkpoint_encoder=True, keep_chunks=16).cuda()
learner = RLLearner(model)
data = tree_map(lambda t: t.to(dev), synth_rl_batch(128, 32, seed=0))
learner._train(data); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    learner._train(data); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ('aten::mm', 'aten::addmm', 'aten::bmm', 'aten::add', 'aten::add_', 'aten::mul', 'aten::copy_', 'aten::sum', 'aten::cat',
                 'aten::native_layer_norm', 'aten::native_layer_norm_backward', 'aten::fill_', 'aten::zero_'):
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
for t, n, k, s in sorted(rows, reverse=True)[:40]:
    print('%9.0f us n=%5d %-28s %s' % (t, n, k, s))
