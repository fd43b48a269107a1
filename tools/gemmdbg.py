import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, bench
d = bench.kernel_rooflines(torch.device('cuda', 0), bench.load_peaks())
print(os.environ.get('DSB_GEMM_DEBUG', '0'), ' '.join('%s=%.0fus' % (k.replace('entity_mlp_gemm_', ''), v['us_per_launch']) for k, v in d.items() if 'M524288' not in k and 'scatter' not in k))
