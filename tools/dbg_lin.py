import sys, math, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from distar_b200 import ops
DEV = 'cuda'
M, K, N, relu, dtype = 264, 90, 128, True, torch.int16
for dtype in (torch.int16, torch.uint8, torch.float32):
  for relu in (True, False):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g) if dtype == torch.float32 else torch.randint(0, 2, (M, K), generator=g).to(dtype)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    go = torch.randn(M, N, generator=g)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = F.linear(x.double(), wr, br)
    ref = torch.relu(ref) if relu else ref
    ref.backward(go.double())
    wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    out = ops.linear(x.to(DEV), wd, bd, relu, 3, exact_input=dtype != torch.float32)
    out.backward(go.to(DEV))
    e = (wd.grad.cpu().double() - wr.grad).abs()
    bad = (e > 1e-3).nonzero()
    print(dtype, relu, 'fwd err', (out.detach().cpu().double() - ref.detach()).abs().max().item(), 'dW err', e.max().item(), 'nbad', bad.shape[0],
          'rows', sorted(set(bad[:, 0].tolist()))[:6], 'cols', sorted(set(bad[:, 1].tolist()))[:12],
          'db err', (bd.grad.cpu().double() - br.grad).abs().max().item())
