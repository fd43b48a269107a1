import sys, torch
sys.path.insert(0, '.')
from distar_b200 import ops
DEV = 'cuda'
g = torch.Generator().manual_seed(1)
for M, Np, Kp, exact in [(264, 128, 128, True), (264, 128, 128, False), (320, 128, 128, True), (264, 128, 64, True), (264, 128, 320, True), (264, 64, 128, True), (200, 128, 128, True), (128, 128, 128, True)]:
    gy = torch.randn(M, Np, generator=g).to(DEV)
    x = torch.randint(0, 2, (M, Kp), generator=g).float().to(DEV)
    g_hi, g_lo = ops.split_bf16(gy)
    x_hi, x_lo = ops.split_bf16(x)
    kk = (M + 63) // 64 * 64
    part = torch.zeros(Np, Kp, device=DEV)
    ops._gemm_ex(a_hi=g_hi, a_lo=g_lo, b_hi=x_hi, b_lo=None if exact else x_lo, a_mn=1, b_mn=1, alpha=1.0, terms=3, c=part, m=Np, n=Kp, k=kk,
                 batch=1, inner=1, splits=1, c_row_split=0, c_accumulate=1, bn=64 if Kp % 128 else 0, b_exact=1 if exact else 0)
    ref = gy.t() @ x
    err = (part - ref).abs()
    bad = (err > 1e-3 * ref.abs().max()).nonzero()
    print(M, Np, Kp, exact, 'max err', err.max().item(), 'scale', ref.abs().max().item(), 'bad', bad.shape[0],
          'rows', sorted(set(bad[:, 0].tolist()))[:8], 'cols', sorted(set(bad[:, 1].tolist()))[:8])
