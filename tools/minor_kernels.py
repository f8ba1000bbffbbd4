"""Evaluation-side and clustering kernels at the amazon shape, outside any model: ``ssl_predict_mask`` (full_predict + _mask_predict
with the mask taken from the device CSR; both of its kernels), ``ssl_topk`` (k = 40) and ``ssl_kmeans_iter`` (NCL, K = 50).  Live CUDA-event timings
as one JSON line; with ``--ncu`` only the launches (for ``ncu --set full -k regex:"predict_|topk_kernel|kmeans"``).

    python tools/minor_kernels.py [--ncu]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sslrec_b200._lib import check, lib          # noqa: E402
from sslrec_b200.kmeans import KMeansClustering  # noqa: E402
from sslrec_b200.trainer import topk             # noqa: E402

U, I, D, BT, DEG, K_TOP, K_CLUSTER = 76469, 83761, 64, 1024, 12, 40, 50


def exact_order_record(dev, layers=3, n_scored=64):
    import numpy as np
    import scipy.sparse as sp
    from sslrec_b200 import engine as E
    from sslrec_b200.data_handler import normalized_adjacency
    from sslrec_b200.graph import GraphPlan
    from synth_graphs import named_graph
    rows, cols, n_user, n_item = named_graph('amazon', seed=2023)
    r, c, v, n = normalized_adjacency(sp.coo_matrix((np.ones(len(rows), dtype=np.float32), (rows, cols)), shape=(n_user, n_item)))
    plan = GraphPlan(r, c, v, n, dev, side_split=n_user)
    e0 = (torch.rand(n, D, generator=torch.Generator().manual_seed(3)) - 0.5) * 0.02          # xavier-like magnitudes
    x = e0.to(dev)
    total = x.clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(layers):
        x = E.spmm_exact(plan, x)
        total = total + x
    b.record()
    torch.cuda.synchronize()
    ms_layers = a.elapsed_time(b)
    # the reference's operators on the host (data_handler_general_cf.py:53-73 layout: column-sorted, uncoalesced COO)
    order = np.lexsort((r, c))
    coo = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([r[order], c[order]]).astype(np.int64)), torch.from_numpy(v[order]), (n, n))
    xs = [e0]
    for _ in range(layers):
        xs.append(torch.spmm(coo, xs[-1]))
    ref = sum(xs)
    got = total.cpu()
    users = torch.arange(0, n_user, n_user // n_scored)[:n_scored]
    preds = torch.empty(n_scored, n_item, device=dev)
    ue_, ie_ = total[:n_user], total[n_user:]
    users_dev = users.to(dev)
    check(lib.ssl_predict_mask(ue_.data_ptr(), ue_.stride(0), ie_.data_ptr(), ie_.stride(0), users_dev.data_ptr(), n_scored, n_item, D, None, None, None,
                               preds.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), 'ssl_predict_mask')
    ref_scores = ref[users] @ ref[n_user:].T
    stats = plan.stats()
    return {'graph': f'synthetic amazon shape: {n_user} x {n_item}, {len(r)} stored entries, longest row {stats["max_row_nnz"]}', 'layers': layers,
            'spmm_exact_ms_per_layer': ms_layers / layers,
            'embeddings_bit_equal_to_reference_cpu_operators': float((got == ref).float().mean().item()),
            'scores_bit_equal_to_reference_cpu_operators': float((preds.cpu() == ref_scores).float().mean().item()),
            'top40_identical': float((topk(preds, K_TOP).cpu() == torch.topk(ref_scores, K_TOP).indices).float().mean().item()),
            'how': 'GPU: engine.spmm_exact per layer + torch adds in the order ((E0 + X1) + X2) + X3, ssl_predict_mask (tiled kernel); host: torch.spmm on the '
                   'reference-layout COO, sum(), matmul -- the operators the reference runs (lightgcn.py:29-43,64)'}


def main():
    ncu = '--ncu' in sys.argv
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    ue = torch.randn(U, D, device=dev, generator=g) * 0.1
    ie = torch.randn(I, D, device=dev, generator=g) * 0.1
    users = torch.randint(0, U, (BT,), device=dev, generator=g)
    # training CSR: DEG distinct items per user, sorted within the row (what BaseModel._train_csr builds from trn_mat)
    u = torch.arange(U, device=dev).unsqueeze(1)
    j = torch.arange(DEG, device=dev).unsqueeze(0)
    cols = ((u * 7 + j * 6997) % I).sort(1).values.to(torch.int32).reshape(-1).contiguous()
    rowptr = (torch.arange(U + 1, device=dev) * DEG).to(torch.int32)
    preds = torch.empty(BT, I, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def predict():
        check(lib.ssl_predict_mask(ue.data_ptr(), ue.stride(0), ie.data_ptr(), ie.stride(0), users.data_ptr(), BT, I, D,
                                   None, rowptr.data_ptr(), cols.data_ptr(), preds.data_ptr(), stream), 'ssl_predict_mask')

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if ncu:
        for _ in range(3):
            predict()
            topk(preds, K_TOP)
        KMeansClustering(K_CLUSTER, D, iters=4, check_every=100)(ue)
        torch.cuda.synchronize()
        return
    out = {'shape': dict(users=U, items=I, dim=D, eval_batch=BT, k=K_TOP, clusters=K_CLUSTER),
           'how': 'CUDA events around 20 back-to-back launches after one warm-up launch; amazon shape, random tables, training CSR of 12 items per user'}
    # both score kernels behind ssl_predict_mask: the 128 x 128 tiled product (default) and the round-1 warp-per-item kernel
    res = {}
    for key, flag in (('predict_mask_warp_per_item', 0), ('predict_mask', 1)):
        check(lib.ssl_set_option(b'predict_tiled', flag), 'ssl_set_option')
        ms = timed(predict, 20)
        res[key] = preds.clone()
        out[key] = {'kernel': 'predict_tile_kernel' if flag else 'predict_mask_kernel', 'ms': ms, 'tflops_fp32': 2.0 * BT * I * D / ms / 1e9,
                    'write_GBps': 4.0 * BT * I / ms / 1e6, 'write_roofline_ms': 4.0 * BT * I / 6490.5e6,
                    'frac_of_hbm_write_roofline': (4.0 * BT * I / 6490.5e6) / ms, 'frac_of_fp32_fma_peak': 2.0 * BT * I * D / ms / 1e9 / 72.3}
    a, b = res['predict_mask'], res['predict_mask_warp_per_item']
    ref = (ue[users].double() @ ie.double().T)
    unmasked = a > -1e7
    out['predict_parity'] = {'masked_positions_equal': bool(torch.equal(a <= -1e7, b <= -1e7)), 'masked_per_row': float((~unmasked).sum().item()) / BT,
                             'tiled_vs_float64_max_abs': float((a.double() - ref)[unmasked].abs().max().item()),
                             'warp_vs_float64_max_abs': float((b.double() - ref)[b > -1e7].abs().max().item()),
                             'top40_identical_between_kernels': float((topk(a, K_TOP) == topk(b, K_TOP)).float().mean().item())}
    # the reference's operator on the host (lightgcn.py:64: users' rows @ item_embeds.T, torch CPU fp32) for 64 of the batch's users: its GEMM evaluates a
    # score as one sequential FMA chain over k, the order of the tiled kernel, so the unmasked scores should agree BIT FOR BIT
    sub = torch.arange(0, BT, BT // 64, device=dev)[:64]
    host = ue[users[sub]].cpu() @ ie.cpu().T
    for key, dev_scores in (('tiled', a), ('warp_per_item', b)):
        got = dev_scores[sub].cpu()
        keep = got > -1e7
        out['predict_parity'][key + '_bit_equal_to_reference_cpu_gemm'] = float((got[keep] == host[keep]).float().mean().item())
    del res, a, b, ref, unmasked, host
    idx = topk(preds, K_TOP)
    ref = torch.topk(preds, K_TOP).indices
    out['topk_matches_torch'] = float((idx == ref).float().mean().item())
    ms = timed(lambda: topk(preds, K_TOP), 20)
    out['topk'] = {'ms': ms, 'read_GBps_one_pass': 4.0 * BT * I / ms / 1e6}
    ms_t = timed(lambda: torch.topk(preds, K_TOP), 20)
    out['torch_topk_ms'] = ms_t
    init = torch.rand(K_CLUSTER, D, generator=g, device=dev)
    cents = {}
    for key, rows in (('kmeans_iter_1_row_per_round', 1), ('kmeans_iter', 4)):      # kmeans_assign_kernel<1> (round-1 form) and <4> (default)
        check(lib.ssl_set_option(b'kmeans_rows_per_round', rows), 'ssl_set_option')
        km = KMeansClustering(K_CLUSTER, D, iters=32, check_every=1000)
        km.init_centroids = init
        km(ue)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = km(ue)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / km.last_iters
        cents[key] = res
        out[key] = {'ms': ms, 'iters': km.last_iters, 'tflops_fp32': 3.0 * U * K_CLUSTER * D / ms / 1e9, 'table_read_GBps': 4.0 * U * D / ms / 1e6}
    out['kmeans_bit_identical'] = bool(all(torch.equal(p, q) for p, q in zip(cents['kmeans_iter'], cents['kmeans_iter_1_row_per_round'])))
    # ---- opt-in evaluation mode test.exact_order: forward pass (engine.spmm_exact per layer, layer sum in the reference's order) + tiled scores on the amazon-shaped
    # synthetic graph against the reference's own operators on this box's CPU (torch.spmm on the column-sorted COO, sum(), @): expected bit-identical ----
    try:
        out['exact_order'] = exact_order_record(dev)
    except Exception as e:      # noqa: BLE001 -- never at the price of the records above
        out['exact_order'] = {'error': repr(e)[:400]}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
