"""The oracle (oracle/cf_oracle.py) against golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only.  Tolerances: the oracle restates the reference's ops with the
same torch CPU kernels, so float32 results agree to a few ulp; the float64 replay bounds the
float32 rounding budget that the CUDA parity tests use."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from oracle import inputs, replay

CASES = [('lightgcn', 'tiny'), ('simgcl', 'tiny'), ('sgl', 'tiny'), ('sgl_nd', 'tiny'), ('ncl', 'tiny'),
         ('hccf', 'tiny'), ('lightgcn', 'small'), ('simgcl', 'small'), ('sgl', 'small'), ('simgcl', 'mid'),
         ('directau', 'tiny'), ('directau', 'small'), ('lightgcl', 'tiny'), ('lightgcl', 'small'),
         ('ncl_k50', 'small'), ('hccf_h128', 'small')]        # the YAML sizes: ncl.yml cluster_num 50, hccf.yml hyper_num 128


def _close(a, b, rtol, atol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b) + (2e-6 * np.abs(b).max() if b.size > 1 else 0.0)   # + fp32 noise of the largest entry
    assert (err <= tol).all(), f'{what}: max err {err.max():.3e} (tol {tol.flat[err.argmax()]:.3e})'


def test_adjacency_matches_reference():
    g = replay.load_golden('lightgcn', 'tiny')
    case = inputs.make_case('tiny')
    adj = O.normalized_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    p = O.coo_order_like_reference(adj)
    assert np.array_equal(adj.rows[p], g['adj_rows'])          # entry order of _make_torch_adj
    assert np.array_equal(adj.cols[p], g['adj_cols'])
    assert np.array_equal(adj.vals[p].view(np.uint32), g['adj_vals'].view(np.uint32))   # bit-exact values
    # symmetric structure and values: one CSR serves forward and backward
    key = adj.rows * adj.n + adj.cols
    keyt = adj.cols * adj.n + adj.rows
    assert np.array_equal(np.sort(key), np.sort(keyt))
    assert np.array_equal(adj.vals[np.argsort(key)].view(np.uint32), adj.vals[np.argsort(keyt)].view(np.uint32))


@pytest.mark.parametrize('model_key,case_name', CASES)
def test_oracle_reproduces_reference(model_key, case_name):
    g = replay.load_golden(model_key, case_name)
    o = replay.oracle_outputs(model_key, case_name, torch.float32, g)
    _close(o['loss'], g['loss'], 2e-6, 1e-7, 'loss')
    for k in g:
        if k.startswith('part_'):
            _close(o[k], g[k], 2e-6, 1e-9, k)
    if 'adj_vals_sum' in g:
        _close(o['adj'].vals.astype(np.float64).sum(), g['adj_vals_sum'], 1e-12, 0, 'adj_vals_sum')
    for k in g:
        if k.startswith('grad_'):
            if k.endswith('_head'):
                name = k[:-5]
                _close(o[name][:32], g[k], 1e-4, 1e-9, k)
            elif k.endswith('_rowsum'):
                name = k[:-7]
                _close(o[name].double().sum(1), g[k], 1e-4, 1e-8, k)
            elif k.endswith('_abssum'):
                name = k[:-7]
                _close(o[name].double().abs().sum(), g[k], 1e-5, 0, k)
            else:
                _close(o[k], g[k], 1e-4, 1e-9, k)
        if k.startswith('new_'):
            got = o[k[:-5]][:32] if k.endswith('_head') else o[k]
            gref = g['grad_' + k[4:]] if 'grad_' + k[4:] in g else g['grad_' + k[4:-5]][:32]     # small tensors keep the full gradient
            pname = k[4:-5] if k.endswith('_head') else k[4:]
            if pname in ('user_embeds', 'item_embeds'):          # Adam folds weight decay (directau.yml: 1e-6) into g
                p0 = o['case'][{'user_embeds': 'user_e', 'item_embeds': 'item_e'}[pname]].numpy()
                gref = gref + float(g.get('opt_weight_decay', 0.0)) * (p0[:32] if k.endswith('_head') else p0)
            # Adam's first step moves an entry by lr * g / (|g| + 1e-8): where the reference gradient is rounding noise
            # (DirectAU has no regulariser, so rows far from the batch get ~1e-12 gradients) its sign, hence +-lr, is not
            # defined; those entries only have to stay within 2 lr
            noise = np.abs(gref) <= 1e-5 * np.abs(gref).max()
            _close(np.where(noise, g[k], np.asarray(got)), g[k], 1e-6, 1e-7, k)
            assert (np.abs(np.asarray(got) - g[k])[noise] <= 2.1e-3).all(), k
    if 'preds' in g:
        _close(o['preds'], g['preds'], 1e-5, 1e-6, 'preds')
    # top-K: identical indices wherever the reference's own score gap exceeds fp32 reassociation noise
    gv, gi = g['topk_val'], g['topk_idx']
    oi = o['topk_idx'].numpy()
    gap_ok = np.ones_like(gi, dtype=bool)
    gap = np.abs(np.diff(gv, axis=1))
    thr = 1e-6 * np.maximum(1.0, np.abs(gv[:, :-1]))
    near = gap <= thr
    gap_ok[:, :-1] &= ~near
    gap_ok[:, 1:] &= ~near
    gap_ok[:, -1] = False       # the K-th place can swap with the (K+1)-th, which the fixture does not hold
    assert (oi[gap_ok] == gi[gap_ok]).all()
    assert gap_ok.mean() > 0.9


@pytest.mark.parametrize('model_key', ['lightgcn', 'simgcl', 'sgl', 'ncl', 'hccf', 'directau', 'lightgcl'])
def test_float64_replay_bounds_fp32_budget(model_key):
    """loss(fp64) - loss(reference fp32) stays inside the 1e-5 budget of BASELINE.json."""
    g = replay.load_golden(model_key, 'tiny')
    o = replay.oracle_outputs(model_key, 'tiny', torch.float64, g)
    assert abs(float(o['loss']) - float(g['loss'])) < 1e-5


def test_kmeans_matches_reference():
    g = replay.load_golden('ncl', 'tiny')
    case = inputs.make_case('tiny')
    adj = O.normalized_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    dr = replay.draws('ncl', case, g['hp'], adj)
    cents, idx, _ = O.kmeans(case['user_e'], dr['init_user_centroids'])
    assert np.array_equal(idx.numpy(), g['user2cluster'])
    _close(cents, g['user_centroids'], 1e-5, 1e-7, 'user_centroids')


def test_negative_sampler_never_returns_a_positive():
    case = inputs.make_case('tiny')
    negs = O.sample_negatives(case['rows'], case['cols'], case['n_item'], np.random.RandomState(3))
    pos = set(zip(case['rows'].tolist(), case['cols'].tolist()))
    assert all((u, j) not in pos for u, j in zip(case['rows'].tolist(), negs.tolist()))
    assert negs.min() >= 0 and negs.max() < case['n_item']


def test_philox_known_answers_and_sampler_properties():
    """Random123's published Philox4x32-10 known-answer vectors pin the numpy generator that the GPU tests compare
    the in-kernel draws with; the rejection sampler on top of it never returns a training positive."""
    from oracle import philox as P
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = P.philox4x32_10(*[np.array([c]) for c in ctr], key[0] | (key[1] << 32))
        assert tuple(int(g[0]) for g in got) == want
    import scipy.sparse as sp
    rows, cols = inputs.bipartite_edges(40, 9, 250, 3)
    csr = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(40, 9))
    csr.sort_indices()
    coo = csr.tocoo()
    negs = P.sample_negs(coo.row, csr.indptr, csr.indices, 9, seed=5, epoch=0)
    pos = set(zip(coo.row.tolist(), coo.col.tolist()))
    full = set(np.flatnonzero(np.diff(csr.indptr) == 9).tolist())           # users who interacted with everything
    assert all((int(u), int(j)) not in pos for u, j in zip(coo.row, negs) if int(u) not in full)
    assert negs.min() >= 0 and negs.max() < 9
    assert not np.array_equal(negs, P.sample_negs(coo.row, csr.indptr, csr.indices, 9, seed=5, epoch=1))
    u = P.noise_uniform(11, 1, 64, 10)
    assert u.shape == (64, 10) and u.dtype == np.float32 and 0.0 <= u.min() and u.max() < 1.0
    assert abs(P.edge_keep(3, 0, np.arange(20000) % 500, np.arange(20000) // 500, 0.25).mean() - 0.25) < 0.02


@pytest.mark.parametrize('case_name', ['tiny', 'small'])
def test_lightgcl_adjacency_matches_reference_bits(case_name):
    """R / sqrt(rowD colD) in float32 (lightgcl.py:16-20) as the oracle lays it out: bit-identical to the reference's tensor."""
    g = replay.load_golden('lightgcl', case_name)
    case = inputs.make_case(case_name)
    la = O.lightgcl_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    upper = la.rows < la.n_user
    og = np.lexsort((g['lgcl_cols'], g['lgcl_rows']))
    assert np.array_equal(la.rows[upper], g['lgcl_rows'][og]) and np.array_equal(la.cols[upper] - la.n_user, g['lgcl_cols'][og])
    assert np.array_equal(la.vals[upper].view(np.uint32), g['lgcl_vals'][og].view(np.uint32))
    # the lower block is the transpose: one symmetric CSR serves _spmm(adj, E_i) and _spmm(adj^T, E_u)
    key, keyt = la.rows * la.n + la.cols, la.cols * la.n + la.rows
    assert np.array_equal(la.vals[np.argsort(key)].view(np.uint32), la.vals[np.argsort(keyt)].view(np.uint32))
