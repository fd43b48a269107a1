"""Dev helper: A/B timing of the FFN-shaped GEMM (M=135168, K=256, N=1024, 3-term) per tile mode and output form.
usage: python tools/gemmab.py [reps]   (CUDA-graph replay, us per launch; checks pair vs single results agree)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200 import lib, ops
dev = torch.device('cuda', 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
modes = [tuple(int(v) for v in m.split(':')) for m in sys.argv[3].split(',')] if len(sys.argv) > 3 else ((0, 1), (256, 4), (128, 4))
check = not os.environ.get('DSB_GEMM_DEBUG')
s, e = torch.cuda.Event(True), torch.cuda.Event(True)


def timed(run):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / reps * 1e3)
    return best


for (M, K, N, tag) in ((264 * 512, 256, 1024, 'ffn1'), (264 * 512, 1024, 256, 'ffn2'), (264 * 512, 256, 768, 'qkv'),
                       (264 * 512, 256, 256, 'proj'), (264 * 256, 1152, 128, 'conv16'), (1024 * 1024, 1152, 128, 'conv32')):
    if only and tag not in only:
        continue
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
    a_hi, a_lo = ops.split_bf16(a); w_hi, w_lo = ops.split_bf16(w)
    c = torch.empty(M, N, device=dev)
    c_hi = torch.empty(M, N, device=dev, dtype=torch.bfloat16); c_lo = torch.empty_like(c_hi)
    ref = None
    row = []
    for out in ('c', 'pair'):
        for bn, mc in modes:
            if N % 256 and bn == 256:
                continue
            kw = dict(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b, alpha=1.0, relu=1, terms=3, m=M, n=N, k=K, batch=1,
                      inner=1, splits=1, bn=bn, mc=mc)
            if out == 'c':
                kw['c'] = c
            else:
                kw.update(c_hi=c_hi, c_lo=c_lo, c_rows=M, c_cols=N)
            run = lambda: lib.gemm_ex(**kw)
            us = timed(run)
            if out == 'c' and check:
                if ref is None:
                    ref = c.clone()
                else:
                    assert torch.equal(ref, c) or (ref - c).abs().max() < 1e-4 * ref.abs().max(), 'pair != single'
            row.append('%s bn%d mc%d %.1f' % (out, bn, mc, us))
    print('dbg', os.environ.get('DSB_GEMM_DEBUG', '0'), tag, M, K, N, '|', ' | '.join(row), flush=True)
