"""Dev helper: device time of every non-GEMM C-ABI call of one learner step, grouped by (entry point, integer arguments)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200 import lib
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
B, T = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0,
              encoder_chunk=264, checkpoint_encoder=True, keep_chunks=16).cuda()
learner = RLLearner(model)
data = tree_map(lambda t: t.to(dev), synth_rl_batch(B, T, seed=0))
learner._train(data); torch.cuda.synchronize()
rec = []
orig_call = lib.call
def wrapped_call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig_call(name, *a); e1.record()
    ints = tuple(x for x in a if isinstance(x, int) and not isinstance(x, bool))
    nbytes = sum(x.numel() * x.element_size() for x in a if isinstance(x, torch.Tensor))
    rec.append(((name, ints), e0, e1, nbytes))
    return r
lib.call = wrapped_call
learner._train(data); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0])
for key, e0, e1, nb in rec:
    a = agg[key]; a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += nb
print('calls', len(rec), 'total ms', sum(v[1] for v in agg.values()))
print('      ms    n   GB/s(tensor args)  call')
for key, (n, t, nb) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print('%8.2f %4d %8.0f  %s %s' % (t, n, nb / (t * 1e-3) / 1e9, key[0].replace('dsb_', ''), key[1]))
