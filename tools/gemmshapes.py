"""Dev helper: per-shape device time of every dsb_gemm_ex launch in one learner step (CUDA events around each call)."""
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
orig = lib.gemm_ex
rec = []
def wrapped(**kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(**kw); e1.record()
    key = (kw.get('m'), kw.get('n'), kw.get('k'), kw.get('batch', 1), kw.get('splits', 1), kw.get('terms', 3), kw.get('bn', 0),
           int(bool(kw.get('a_conv', 0))) + 2 * int(bool(kw.get('b_conv', 0))), kw.get('a_mn', 0), kw.get('b_mn', 0),
           int(kw.get('c_hi') is not None), int(bool(kw.get('c_accumulate', 0))))
    rec.append((key, e0, e1))
lib.gemm_ex = wrapped
orig_call = lib.call
def wrapped_call(name, *a):
    if name != 'dsb_gemm_bf16_split':
        return orig_call(name, *a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig_call(name, *a); e1.record()
    rec.append((('fwd', a[8], a[9], a[10], 1, 1, a[11], 0, 0, 0, 0, int(a[6] is not None), 0)[1:], e0, e1))
lib.call = wrapped_call
import distar_b200.ops as ops
learner._train(data); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in rec:
    agg[key][0] += 1; agg[key][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print('gemm launches', len(rec), 'total ms', tot)
print('      ms   n    TF   (m, n, k, batch, splits, terms, bn, conv, a_mn, b_mn, emit_pair, accumulate)')
for key, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    m, nn, k, b, sp, terms = key[:6]
    fl = 2.0 * m * nn * k * b * terms * n
    print('%8.2f %4d %6.0f  %s' % (t, n, fl / (t * 1e-3) / 1e12, key))
