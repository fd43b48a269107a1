"""Dev helper: time of every tcgen05 GEMM launch of one learner step (lib.GEMM_TRACE), grouped by problem shape."""
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
for _ in range(2):
    learner._train(data)
torch.cuda.synchronize()
shapes = []
orig = lib.gemm_ex
def traced(**kw):
    shapes.append((kw['m'], kw['n'], kw['k'], kw.get('batch', 1), kw.get('splits', 1), kw.get('a_mn', 0), kw.get('b_mn', 0),
                   1 if kw.get('a_conv') or kw.get('b_conv') else 0, 'pair' if kw.get('c') is None else ('c+pair' if kw.get('c_hi') is not None else 'c')))
    orig(**kw)
lib.gemm_ex = traced
import distar_b200.ops as ops
ops._gemm_ex = lambda **kw: traced(**kw)
orig_call = lib.call
def call(name, *a):
    if name == 'dsb_gemm_bf16_split':
        shapes.append((a[8], a[9], a[10], 1, 1, 0, 0, 0, 'c' if a[6] is None else ('pair' if a[5] is None else 'c+pair')))
    return orig_call(name, *a)
lib.call = call
ops.lib.call = call
lib.GEMM_TRACE = []
learner._train(data)
torch.cuda.synchronize()
tr = lib.GEMM_TRACE
lib.GEMM_TRACE = None
assert len(tr) == len(shapes), (len(tr), len(shapes))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (e0, e1, fl, pr), sh in zip(tr, shapes):
    a = agg[sh]; a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl * pr
tot = sum(v[1] for v in agg.values())
print('total gemm ms %.1f launches %d' % (tot, len(tr)))
print('%10s %6s %6s %6s %4s %2s%2s %4s %7s | %5s %9s %7s %8s' % ('m', 'n', 'k', 'batch', 'spl', 'a', 'b', 'conv', 'out', 'n', 'ms', 'us/call', 'TF(work)'))
for sh, (n, ms, work) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print('%10d %6d %6d %6d %4d %2d%2d %4d %7s | %5d %9.2f %7.1f %8.0f' % (*sh, n, ms, 1e3 * ms / n, work / ms / 1e9))
