import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
g = torch.Generator().manual_seed(0)
mode = os.environ.get('SSL_TC_DEBUG', '0')
R = (torch.randint(-4, 5, (nr, d), generator=g).float() / 8).cuda()      # tf32-exact small values
C = (torch.randint(-4, 5, (nc, d), generator=g).float() / 8).cuda()
rs, o = call(R, C)
S = R.double().cpu() @ C.double().cpu().T
E = torch.exp2(S)
if mode == '1':
    exp = R.double().cpu() @ C.double().cpu()
elif mode == '2':
    Ehi = (E.float().view(torch.int32) & -8192).view(torch.float32).double()
    exp = Ehi @ C.double().cpu().T
else:
    exp = E @ C.double().cpu()
print('mode', mode, 'rowsum err', (rs.double() - E.sum(1)).abs().max().item() / E.sum(1).abs().max().item())
print('O max abs', o.abs().max().item(), 'exp max abs', exp.abs().max().item(), 'max err', (o.double() - exp).abs().max().item())
print('got ', o[0][:8].tolist()); print('exp ', exp[0][:8].tolist())
print('got r5', o[5][:8].tolist()); print('exp r5', exp[5][:8].tolist())
