"""Kernel-level parity on the GPU, through the C ABI (ctypes): each kernel against the oracle /
a float64 torch restatement of the same formula on seeded inputs, plus size-independent
properties (adjointness of the masked SpMM, keep fractions, noise norms, determinism)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from oracle import inputs
import ssl_test_helpers as H

pytestmark = pytest.mark.gpu


def _graph(n_user, n_item, n_edge, seed, hub=0):
    rows, cols = inputs.bipartite_edges(n_user, n_item, n_edge, seed)
    if hub:      # one item connected to `hub` users: exercises the split-row path (> 128 entries)
        extra_u = np.arange(hub) % n_user
        rows = np.concatenate([rows, extra_u])
        cols = np.concatenate([cols, np.full(hub, n_item - 1)])
    return O.normalized_adjacency(rows, cols, n_user, n_item)


def _plan(adj, need_rev=False):
    from sslrec_b200.graph import GraphPlan
    return GraphPlan(adj.rows, adj.cols, adj.vals, adj.n, torch.device('cuda'), need_rev=need_rev)


@pytest.mark.parametrize('dim', [16, 32, 64, 128, 48])
@pytest.mark.parametrize('hub', [0, 700])
def test_propagation_matches_oracle(dim, hub):
    from sslrec_b200 import engine as E
    adj = _graph(900, 700, 8000, 5, hub)
    plan = _plan(adj)
    if hub:
        assert plan.stats()['split_rows'] >= 1
    g = torch.Generator().manual_seed(1)
    e0 = torch.randn(adj.n, dim, generator=g) * 0.1
    for L in (1, 3):
        prop = E.Propagation(plan, [E.ViewSpec()], L)
        st = prop.forward(e0.cuda(), 900)
        ref = O.lightgcn_embeds(adj.torch_coo(torch.float64), e0.double(), L)
        H.close(st.E.view(adj.n, dim), ref, 1e-5, 1e-6, f'E L={L}')
        again = prop.forward(e0.cuda(), 900)
        assert torch.equal(st.E, again.E)              # fixed summation order -> bit-reproducible


def test_three_views_share_layer_one_and_match_single_views():
    from sslrec_b200 import engine as E
    adj = _graph(500, 400, 5000, 6, 300)
    plan = _plan(adj)
    e0 = (torch.randn(adj.n, 64, generator=torch.Generator().manual_seed(2)) * 0.1).cuda()
    views = [E.ViewSpec(noise_mode=1, seed=11), E.ViewSpec(noise_mode=1, seed=12), E.ViewSpec()]
    st3 = E.Propagation(plan, views, 3, noise_eps=0.2).forward(e0, 500)
    for v, spec in enumerate(views):
        st1 = E.Propagation(plan, [spec], 3, noise_eps=0.2).forward(e0, 500)
        assert torch.equal(st3.E[:, v, :], st1.E[:, 0, :])


def test_injected_noise_matches_oracle_and_rng_noise_has_norm_eps():
    from sslrec_b200 import engine as E
    adj = _graph(300, 200, 3000, 7)
    plan = _plan(adj)
    g = torch.Generator().manual_seed(3)
    e0 = torch.randn(adj.n, 32, generator=g) * 0.1
    us = [torch.rand(adj.n, 32, generator=g) for _ in range(2)]
    st = E.Propagation(plan, [E.ViewSpec(noise_mode=2, noise_u=[u.cuda() for u in us])], 2, noise_eps=0.9).forward(e0.cuda(), 300)
    ref = O.simgcl_embeds(adj.torch_coo(torch.float64), e0.double(), 2, 0.9, [u.double() for u in us])
    H.close(st.E.view(adj.n, 32), ref, 1e-5, 1e-6, 'perturbed E')
    # RNG noise: x_out - A x has row norm eps wherever no entry of the clean output is exactly 0
    clean = E.Propagation(plan, [E.ViewSpec()], 1, sum_layers=1, keep_layers=(1,)).forward(e0.cuda(), 300).layers[1]
    noisy = E.Propagation(plan, [E.ViewSpec(noise_mode=1, seed=99)], 1, sum_layers=1, keep_layers=(1,), noise_eps=0.9).forward(e0.cuda(), 300).layers[1]
    diff = (noisy - clean).view(adj.n, 32)
    full = (clean.view(adj.n, 32) != 0).all(1)
    norms = diff[full].norm(dim=1)
    assert full.sum() > 100 and torch.allclose(norms, torch.full_like(norms, 0.9), rtol=1e-4)
    assert (torch.sign(diff[full]) == torch.sign(clean.view(adj.n, 32)[full])).all()       # noise follows sign(X)
    assert (diff[~full][clean.view(adj.n, 32)[~full] == 0] == 0).all()


@pytest.mark.parametrize('mode', ['rng', 'injected'])
def test_masked_spmm_is_adjoint_of_its_transpose(mode):
    """<A_m x, y> == <x, A_m^T y> with the edge mask evaluated in-kernel: the backward kernel
    (transpose = 1, key swapped / rev-indexed) is the exact transpose of the forward one."""
    from sslrec_b200 import engine as E
    adj = _graph(800, 600, 9000, 8, 400)
    plan = _plan(adj, need_rev=True)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(adj.n, 64, generator=g).cuda()
    y = torch.randn(adj.n, 64, generator=g).cuda()
    keep = 0.5
    if mode == 'rng':
        view = E.ViewSpec(edge_mode=1, keep=keep, scale=2.0, seed=1234)
    else:
        m = (torch.rand(adj.nnz, generator=g) < keep).to(torch.uint8).cuda()
        view = E.ViewSpec(edge_mode=2, keep=keep, scale=2.0, edge_masks=m)
    prop = E.Propagation(plan, [view], 1)

    def apply(v, transpose):
        a = prop._args(64, 1, transpose)
        out = torch.empty(adj.n, 1, 64, device='cuda')
        a.in_views, a.x_in, a.x_out = 1, v.data_ptr(), out.data_ptr()
        prop._launch(a, v)
        return out.view(adj.n, 64)
    ax, aty = apply(x, False), apply(y, True)
    lhs, rhs = (ax.double() * y.double()).sum().item(), (x.double() * aty.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0)
    # keep fraction of the RNG mask ~ keep, and it is asymmetric (each directed entry drawn independently)
    ones = torch.ones(adj.n, 64, device='cuda')
    kept = apply(ones, False)[:, 0].double().sum().item() / (2.0 * float(adj.vals.astype(np.float64).sum()))
    assert abs(kept - keep) < 0.02
    if mode == 'injected':
        ref = torch.spmm(O.edged(adj, m) if False else O.edge_dropped(adj, m.cpu().numpy().astype(bool), keep, True, torch.float64), x.cpu().double())
        H.close(ax, ref, 1e-5, 1e-6, 'injected-mask SpMM')


def test_node_drop_forward_backward():
    from sslrec_b200 import engine as E
    adj = _graph(300, 200, 3000, 9)
    plan = _plan(adj)
    g = torch.Generator().manual_seed(5)
    e0 = torch.randn(adj.n, 32, generator=g) * 0.1
    mask = (torch.rand(adj.n, generator=g) < 0.5)
    view = E.ViewSpec(node_mode=2, node_keep=0.5, node_mask=mask.to(torch.uint8).cuda())
    prop = E.Propagation(plan, [view, E.ViewSpec()], 2)
    st = prop.forward(e0.cuda(), 300)
    a_t = adj.torch_coo(torch.float64)
    ref0 = O.lightgcn_embeds(a_t, O.node_dropped(e0.double(), mask), 2)
    ref1 = O.lightgcn_embeds(a_t, e0.double(), 2)
    H.close(st.E[:, 0, :], ref0, 1e-5, 1e-6, 'node-dropped view')
    H.close(st.E[:, 1, :], ref1, 1e-5, 1e-6, 'clean view')
    # backward: d/dE0 of sum(E * W)
    w = torch.randn(adj.n, 2, 32, generator=g)
    st.g_sum().copy_(w.cuda())
    de0 = prop.backward(st)
    e0r = e0.double().clone().requires_grad_(True)
    tot = (O.lightgcn_embeds(a_t, O.node_dropped(e0r, mask), 2) * w[:, 0].double()).sum() + (O.lightgcn_embeds(a_t, e0r, 2) * w[:, 1].double()).sum()
    tot.backward()
    H.close(de0, e0r.grad, 1e-5, 1e-5, 'dE0 with node drop')


@pytest.mark.parametrize('use_tc', [True, False])
@pytest.mark.parametrize('dim,B,n', [(64, 4096, 9000), (32, 100, 777), (128, 300, 2000), (48, 257, 1000), (64, 64, 50), (64, 130, 64 * 9 + 1)])
def test_infonce_term_forward_backward(dim, B, n, use_tc, monkeypatch):
    """use_tc: the tcgen05 3xTF32 contraction (dims 32 / 64) vs the FP32-FMA kernel -- same tolerances."""
    from sslrec_b200 import engine
    from sslrec_b200 import loss_utils as LU
    monkeypatch.setattr(engine, 'USE_TENSOR_CORES', use_tc)
    g = torch.Generator().manual_seed(6)
    e1 = torch.randn(B, dim, generator=g)
    e2 = torch.randn(B, dim, generator=g)
    tab = torch.randn(n, dim, generator=g)
    tau = 0.2
    ins = [t.clone().cuda().requires_grad_(True) for t in (e1, e2, tab)]
    loss = LU.cal_infonce_loss(*ins, tau)
    loss.backward()
    ref_in = [t.double().clone().requires_grad_(True) for t in (e1, e2, tab)]
    ref = O.infonce_loss_sum(*ref_in, tau)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    # absolute term relative to the largest gradient entry: 2e-6 for the FP32-FMA kernel; 1e-5 for the tcgen05
    # 3xTF32 kernel, whose tensor-core accumulators round toward zero over up to ~10^3 accumulations per output
    # (measured 0.5-2.5e-6 of the largest entry, tools/debug_tc.py)
    rel_atol = 1e-5 if (use_tc and dim in (32, 64)) else 2e-6
    for a, b, name in zip(ins, ref_in, ('e1', 'e2', 'table')):
        H.close(a.grad, b.grad, 2e-4, rel_atol * b.grad.abs().max().item(), 'grad ' + name)


def test_spec_nodes_infonce_and_bpr_dense():
    from sslrec_b200 import loss_utils as LU
    g = torch.Generator().manual_seed(7)
    e1, e2 = torch.randn(500, 32, generator=g), torch.randn(500, 32, generator=g)
    nodes = torch.unique(torch.randint(0, 500, (200,), generator=g))
    a = e2.clone().cuda().requires_grad_(True)
    loss = LU.cal_infonce_loss_spec_nodes(e1.cuda(), a, nodes.cuda(), 0.1)
    loss.backward()
    b = e2.double().clone().requires_grad_(True)
    ref = O.infonce_spec_nodes_mean(e1.double(), b, nodes, 0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-6
    H.close(a.grad, b.grad, 2e-4, 2e-6 * b.grad.abs().max().item(), 'spec-nodes grad')
    anc, pos, neg = (torch.randn(300, 64, generator=g) for _ in range(3))
    ins = [t.clone().cuda().requires_grad_(True) for t in (anc, pos, neg)]
    l = LU.cal_bpr_loss(*ins)
    l.backward()
    rin = [t.double().clone().requires_grad_(True) for t in (anc, pos, neg)]
    r = O.bpr_loss_sum(*rin)
    r.backward()
    assert abs(l.item() - r.item()) <= 2e-6 * abs(r.item())
    for x, y in zip(ins, rin):
        H.close(x.grad, y.grad, 1e-5, 1e-6, 'bpr grad')


def test_topk_exact_with_ties():
    from sslrec_b200.trainer import topk
    g = torch.Generator().manual_seed(8)
    p = torch.randn(37, 5003, generator=g)
    p[:, 100:140] = 0.5                      # a run of ties
    p[3, :] = -1e8                           # fully masked row
    p[5, 17] = float('inf')
    idx, val = topk(p.cuda(), 40, return_values=True)
    tv, _ = torch.topk(p, 40)
    assert torch.equal(val.cpu(), tv)
    # ties resolve to the lower index, and every returned index carries the returned value
    assert torch.equal(p.gather(1, idx.cpu()), val.cpu())
    srt = torch.sort(torch.stack([-p[0], torch.arange(5003).float()], 1)[:, 0], stable=True).indices[:40]
    assert torch.equal(idx[0].cpu(), srt)


def test_adam_matches_torch():
    from sslrec_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(9)
    w = torch.randn(1003, 33, generator=g)
    a = torch.nn.Parameter(w.clone().cuda())
    b = torch.nn.Parameter(w.clone())
    oa, ob = FusedAdam([a], lr=1e-3, weight_decay=1e-4), torch.optim.Adam([b], lr=1e-3, weight_decay=1e-4)
    for _ in range(4):
        gr = torch.randn(1003, 33, generator=g)
        a.grad, b.grad = gr.cuda(), gr.clone()
        oa.step(); ob.step()
    H.close(a, b, 1e-6, 1e-7, 'adam params')
    H.close(oa.state[a]['exp_avg_sq'], ob.state[b]['exp_avg_sq'], 1e-6, 1e-12, 'adam v')


def test_c_abi_rejects_bad_arguments():
    from sslrec_b200 import _lib
    rc = _lib.lib.ssl_sumsq(None, 4, None, None)
    assert rc == -1 and b'null' in _lib.lib.ssl_last_error()
    x = torch.zeros(8, device='cuda')
    rc = _lib.lib.ssl_rows_normalize(x.data_ptr(), 6, None, 1, 6, 0, 1.0, x.data_ptr(), None, None, None, None, None, None, 0, None)
    assert rc == -1                                               # dim must be a multiple of 4
