"""Dev helper: host-side time of one learner step (how long until every kernel is queued) + cProfile of it."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
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
    learner._train(data); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.time(); learner._train(data); t1 = time.time(); torch.cuda.synchronize(); t2 = time.time()
    print('host queued everything after %.1f ms; device finished after %.1f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
pr = cProfile.Profile()
pr.enable(); learner._train(data); pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(28)
