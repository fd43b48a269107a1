"""Dev helper: forward time per section of the learner step at full size (CUDA events, no_grad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200.model import Model
from distar_b200.synth import synth_rl_batch, tree_map
B, T, rows = 128, 32, 264
dev = torch.device('cuda', 0)
model = Model({'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}, use_value_network=True, seed=0).cuda()
net = model._net()
data = tree_map(lambda t: t.to(dev), synth_rl_batch(B, T, seed=0))
sl = lambda tree: tree_map(lambda t: t[:rows], tree)
sp, en, sc, num = sl(data['spatial_info']), sl(data['entity_info']), sl(data['scalar_info']), data['entity_num'][:rows]
def timeit(name, fn, reps=3):
    with torch.no_grad():
        out = fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(reps): out = fn()
        e.record(); torch.cuda.synchronize()
    print('%-28s %9.2f ms' % (name, s.elapsed_time(e) / reps)); return out
timeit('scalar_encoder[264]', lambda: net.scalar_encoder(sc))
feats = timeit('entity_features[264]', lambda: net.entity_features(en))
ee, emb, mask = timeit('entity_encoder[264]', lambda: net.entity_encoder(en, num))
proj = timeit('scatter_project[264]', lambda: net.fc('encoder.scatter_project', ee, relu=True))
from distar_b200 import ops
smap = timeit('scatter_connection[264]', lambda: ops.scatter_connection(proj, en['x'], en['y'], num, 128, 128))
es, skips = timeit("spatial_encoder[264]", lambda: net.spatial_encoder(sp, proj, en["x"], en["y"], num))
P = 1024
x = torch.randn(33, 128, 1536, device=dev)
st = [(torch.randn(128, 384, device=dev), torch.randn(128, 384, device=dev)) for _ in range(3)]
timeit('core_lstm[33x128]', lambda: net.lstm('core_lstm', x, st, 3))
lo = torch.randn(P, 384, device=dev); ctx = torch.randn(P, 448, device=dev)
at = torch.randint(0, 327, (P,), device=dev)
lg, a, embd = timeit('action_type_head[1024]', lambda: net.action_type_head(lo, ctx, at))
ee_p = torch.randn(P, 512, 256, device=dev); enp = torch.full((P,), 512, device=dev)
from distar_b200.synth import synth_actions
act, sunum = synth_actions(P, enp.cpu(), torch.Generator().manual_seed(0), 12)
su = act['selected_units'].to(dev); sunum = sunum.to(dev)
timeit('selected_units_train[1024]', lambda: net.selected_units_train(embd, ee_p, enp, sunum, su))
timeit('target_unit_head[1024]', lambda: net.target_unit_head(embd, ee_p, enp, at % 512))
ms = [None] * 3 + [torch.randn(P, 16, 16, 128, device=dev) for _ in range(4)]
timeit('location_head[1024]', lambda: net.location_head(embd, ms, at))
lo2 = torch.randn(4224, 384, device=dev)
timeit('value_baseline[4224]', lambda: net.value_baseline('winloss', lo2))
