"""Structured-operand probes of ssl_softmax_gemm_tf32x3 (GEMM2 addressing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sslrec_b200._lib import lib, check
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
d, nr, nc = 64, 128, 64
f = dict(device='cuda', dtype=torch.float32)
def call(R, C, off=0.0):
    z_r, z_c = torch.zeros_like(R), torch.zeros_like(C)
    rs, o = torch.zeros(1, nr, **f), torch.zeros(1, nr, d, **f)
    check(lib.ssl_softmax_gemm_tf32x3(R.data_ptr(), z_r.data_ptr(), nr, C.data_ptr(), z_c.data_ptr(), nc, d, None, off, 1, rs.data_ptr(), o.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return rs[0].cpu(), o[0].cpu()
j = torch.arange(nc, **f).view(-1, 1); c = torch.arange(d, **f).view(1, -1)
# T1: E = 1, C[j][c] = (j+1) + (c+1)/128
R = torch.zeros(nr, d, **f); C = (j + 1) + (c + 1) / 128
rs, o = call(R, C)
exp = C.sum(0).cpu()
print('T1 rowsum', rs[:4].tolist(), 'expected 64'); print('T1 O[0] got', o[0][:10].tolist()); print('T1 O[0] exp', exp[:10].tolist()); print('T1 rows identical', bool((o == o[0]).all()), 'max abs err', (o[0] - exp).abs().max().item())
# T2: one-hot row j0
for j0 in (0, 1, 8, 9, 37):
    C = torch.zeros(nc, d, **f); C[j0] = c + 1
    rs, o = call(R, C)
    print(f'T2 j0={j0}: O[0][:12]', o[0][:12].tolist(), ' O[0][30:36]', o[0][30:36].tolist(), 'sum', o[0].sum().item(), 'exp sum', (d * (d + 1) / 2))
# T3: E[r][j] = r+1
C = torch.zeros(nc, d, **f); C[:, 0] = 1.0
R = torch.zeros(nr, d, **f); R[:, 0] = torch.log2(torch.arange(nr, **f) + 1)
rs, o = call(R, C)
print('T3 rowsum[:6]', rs[:6].tolist(), 'exp', [64.0 * (r + 1) for r in range(6)]); print('T3 O[:6,0]', o[:6, 0].tolist(), ' O[100:103,0]', o[100:103, 0].tolist(), 'exp 64(r+1)')
# T4: E[r][j] = j+1
R = torch.zeros(nr, d, **f); R[:, 0] = 1.0
C = torch.zeros(nc, d, **f); C[:, 0] = torch.log2(torch.arange(nc, **f) + 1); C[:, 1] = 1.0; C[:, 2] = torch.arange(nc, **f)
rs, o = call(R, C)
jj = torch.arange(nc, dtype=torch.float64)
print('T4 rowsum[0]', rs[0].item(), 'exp 2080'); print('T4 O[0][:3]', o[0][:3].tolist(), 'exp', [((jj + 1) * torch.log2(jj + 1)).sum().item(), 2080.0, ((jj + 1) * jj).sum().item()])
