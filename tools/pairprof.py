"""Dev helper for ncu: the FFN-shaped GEMM once per mode (single-CTA 128x256, CTA-pair 256x256)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from distar_b200 import lib, ops
dev = torch.device('cuda', 0)
M, K, N = 264 * 512, 256, 1024
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
a_hi, a_lo = ops.split_bf16(a); w_hi, w_lo = ops.split_bf16(w)
c = torch.empty(M, N, device=dev)
for mc in (1, 4, 1, 4):
    lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b, alpha=1.0, relu=1, terms=3, c=c, m=M, n=N, k=K, batch=1,
                inner=1, splits=1, bn=256, mc=mc)
    torch.cuda.synchronize()
