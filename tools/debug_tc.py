"""Compare ssl_softmax_gemm_tf32x3 with ssl_softmax_gemm on small shapes (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sslrec_b200._lib import lib, check

def run(B, n, d, n_split=1, colscale=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, d, generator=g).cuda(); t = torch.randn(n, d, generator=g).cuda()
    Bp, npad = (B + 63) // 64 * 64, (n + 63) // 64 * 64
    f = dict(device='cuda', dtype=torch.float32)
    a_hat, a_t, a_hi, a_lo = torch.empty(Bp, d, **f), torch.empty(Bp // 64, d, 64, **f), torch.empty(Bp, d, **f), torch.empty(Bp, d, **f)
    a_thi, a_tlo, t_thi, t_tlo = torch.empty(d, Bp, **f), torch.empty(d, Bp, **f), torch.empty(d, npad, **f), torch.empty(d, npad, **f)
    t_hat, t_t, t_hi, t_lo = torch.empty(npad, d, **f), torch.empty(npad // 64, d, 64, **f), torch.empty(npad, d, **f), torch.empty(npad, d, **f)
    r1, r2 = torch.empty(B, **f), torch.empty(n, **f)
    s = torch.cuda.current_stream().cuda_stream
    off = 1.4426950408889634 / 0.2
    check(lib.ssl_rows_normalize(x.data_ptr(), d, None, B, d, 0, off, a_hat.data_ptr(), a_t.data_ptr(), r1.data_ptr(), a_hi.data_ptr(), a_lo.data_ptr(), a_thi.data_ptr(), a_tlo.data_ptr(), Bp, s))
    check(lib.ssl_rows_normalize(t.data_ptr(), d, None, n, d, 0, 1.0, t_hat.data_ptr(), t_t.data_ptr(), r2.data_ptr(), t_hi.data_ptr(), t_lo.data_ptr(), t_thi.data_ptr(), t_tlo.data_ptr(), npad, s))
    cs = (torch.rand(npad, generator=g) + 0.5).cuda() if colscale else None
    outs = []
    for tc in (False, True):
        rs, o = torch.zeros(n_split, B, **f), torch.zeros(n_split, B, d, **f)
        if tc:
            check(lib.ssl_softmax_gemm_tf32x3(a_hi.data_ptr(), a_lo.data_ptr(), B, t_hi.data_ptr(), t_lo.data_ptr(), t_thi.data_ptr(), t_tlo.data_ptr(), npad, n, d, None if cs is None else cs.data_ptr(), off, n_split, rs.data_ptr(), o.data_ptr(), s))
        else:
            check(lib.ssl_softmax_gemm(a_hat.data_ptr(), B, t_hat.data_ptr(), t_t.data_ptr(), n, d, None if cs is None else cs.data_ptr(), off, n_split, rs.data_ptr(), o.data_ptr(), s))
        torch.cuda.synchronize()
        outs.append((rs.sum(0).double().cpu(), o.sum(0).double().cpu()))
    # float64 reference
    A, T = a_hat[:B].double().cpu(), t_hat[:n].double().cpu()
    E = torch.exp2(A @ T.T - off)
    if cs is not None: E = E * cs[:n].double().cpu()
    ref_rs, ref_o = E.sum(1), E @ T
    for name, (rs_, o_) in zip(('ffma', 'tc  '), outs):
        er, eo = ((rs_ - ref_rs).abs() / ref_rs.abs()).max().item(), (o_ - ref_o).abs().max().item() / ref_o.abs().max().item()
        blk = [(o_[:, c:c + 16] - ref_o[:, c:c + 16]).abs().max().item() / ref_o.abs().max().item() for c in range(0, d, 16)]
        rb = [(o_[r:r + 32] - ref_o[r:r + 32]).abs().max().item() / ref_o.abs().max().item() for r in range(0, min(B, 128), 32)]
        print(f'B={B} n={n} d={d} split={n_split} cs={colscale} {name}: rowsum rel err {er:.2e}  O rel err {eo:.2e}  per-16col {["%.1e" % b for b in blk]} per-32row {["%.1e" % b for b in rb]}')

for args in [(128, 64, 64), (128, 128, 64), (128, 64 * 5, 64), (128, 64 * 5, 64, 1, True), (256, 1000, 64, 2), (128, 64 * 3, 32), (4096, 9000, 64, 8)]:
    run(*args)
