"""Dev helper: one forward + backward of the fused spatial stem at the learner's chunk size (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200 import ops
from distar_b200.synth import synth_obs, tree_map
N = int(sys.argv[1]) if len(sys.argv) > 1 else 264
obs = tree_map(lambda t: t.cuda(), synth_obs(N, seed=0))
proj = torch.rand(N, 512, 32, device='cuda').requires_grad_(True)
w = (torch.randn(32, 56, 1, 1, device='cuda') / 7).requires_grad_(True)
b = torch.zeros(32, device='cuda').requires_grad_(True)
for _ in range(2):
    out = ops.spatial_stem(obs['spatial_info'], proj, obs['entity_info']['x'], obs['entity_info']['y'], obs['entity_num'], w, b, 64)
    out.sum().backward()
torch.cuda.synchronize()
s, e = torch.cuda.Event(True), torch.cuda.Event(True)
s.record()
out = ops.spatial_stem(obs['spatial_info'], proj, obs['entity_info']['x'], obs['entity_info']['y'], obs['entity_num'], w, b, 64)
e.record(); torch.cuda.synchronize()
print('fwd ms', s.elapsed_time(e))
g = torch.ones_like(out)
s.record(); out.backward(g); e.record(); torch.cuda.synchronize()
print('bwd ms', s.elapsed_time(e))
