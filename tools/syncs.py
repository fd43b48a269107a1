"""Dev helper: list every host<->device synchronisation of one learner step (torch sync debug mode)."""
import os, sys, warnings, traceback, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200.learner import RLLearner
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0,
              encoder_chunk=66, checkpoint_encoder=True, keep_chunks=1).cuda()
learner = RLLearner(model)
data = tree_map(lambda t: t.to(dev), synth_rl_batch(16, 8, seed=0))
learner._train(data); torch.cuda.synchronize()
seen = collections.Counter()
def hook(message, category, filename, lineno, file=None, line=None):
    st = [f for f in traceback.extract_stack() if '/distar_b200/' in f.filename or f.filename.endswith('learner.py')]
    key = ' <- '.join('%s:%d' % (os.path.basename(f.filename), f.lineno) for f in st[-3:])
    seen[key] += 1
warnings.showwarning = hook
warnings.simplefilter('always')
torch.cuda.set_sync_debug_mode('warn')
learner._train(data)
torch.cuda.set_sync_debug_mode('default')
for k, n in seen.most_common():
    print(n, k)
