"""Host-side logic of the autograd-composed drop-ins (HCCF, LightGCL) on CPU: the three native entry points they call are
replaced by the oracle's torch restatements, so the Python composition (layer loop, dropout injection, loss terms, their
order and weights, the parameters that receive gradients) is checked against the reference's golden vectors without a GPU.
The kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from oracle import inputs, replay
import ssl_test_helpers as H


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    assert (err <= atol + rtol * np.abs(b) + 2e-6 * np.abs(b).max()).all(), f'{what}: max err {err.max():.3e}'


def _setup(model_key):
    g = replay.load_golden(model_key, 'tiny')
    case = inputs.make_case('tiny')
    adj = O.normalized_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    dr = replay.draws(model_key, case, g['hp'], adj)
    return g, case, adj, dr


def test_hccf_composition_matches_reference_on_cpu(monkeypatch):
    from sslrec_b200.general_cf import hccf as M
    g, case, adj, dr = _setup('hccf')
    hp = g['hp']
    inject = {'edge_masks_per_layer': [torch.from_numpy(m.astype(np.uint8)) for m in dr['edge_keeps']], 'hyper_keeps': dr['hyper_keeps']}
    model, _ = H.make_model('hccf', case, hp, inject=inject, device='cpu')
    model.load_state_dict({'user_embeds': case['user_e'], 'item_embeds': case['item_e'],
                           'user_hyper_embeds': dr['user_w'], 'item_hyper_embeds': dr['item_w']})

    def gcn_layer(embeds, view, layer):                       # stands in for engine.spmm (ssl_propagate_layer)
        assert view.edge_mode == 2 and abs(view.scale - 1.0 / hp['keep_rate']) < 1e-12
        mask = view.edge_mask_for(layer).numpy().astype(bool)
        return torch.sparse.mm(O.edge_dropped(adj, mask, hp['keep_rate'], True, embeds.dtype), embeds)
    monkeypatch.setattr(model, '_gcn_layer', gcn_layer)
    monkeypatch.setattr(M, 'cal_bpr_loss', O.bpr_loss_sum)
    monkeypatch.setattr(M, 'cal_infonce_loss_spec_nodes', O.infonce_spec_nodes_mean)
    # the hyper-graph kernels (ssl_rowgemm / ssl_colgemm / ssl_hyper_dropout behind engine.hyper_*) -> torch restatements
    from sslrec_b200 import engine as E
    nu = case['n_user']
    drops = []

    def hyper_layer(x, a_u, a_i, slope, drop_u, drop_i):
        outs = []
        for a, drop, xs in ((a_u, drop_u, x[:nu]), (a_i, drop_i, x[nu:])):
            drops.append(drop)
            assert drop.keep == hp['keep_rate'] and drop.mask is not None            # the injected Bernoulli keeps reach the kernel call
            h = a * drop.mask.to(a.dtype) / drop.keep
            outs.append(O.leaky(h @ O.leaky(h.T @ xs, slope), slope))                    # hccf.py:105-106
        return torch.cat(outs, 0)
    monkeypatch.setattr(E, 'hyper_incidence', lambda e, w, mult: e @ w * mult)
    monkeypatch.setattr(E, 'hyper_layer', hyper_layer)
    batch = [torch.from_numpy(case[k]) for k in ('ancs', 'poss', 'negs')]
    loss, parts = model.cal_loss(batch)
    assert len(drops) == 2 * hp['layer_num'] and len({(d.stream) for d in drops}) == len(drops)      # one draw per (layer, side)
    assert list(parts) == ['bpr_loss', 'reg_loss', 'cl_loss']
    _close(loss.item(), g['loss'], 2e-6, 1e-7, 'loss')
    for k, v in parts.items():
        _close(float(v), g['part_' + k], 2e-6, 1e-9, k)
    loss.backward()
    for name, p in model.named_parameters():
        _close(p.grad, g['grad_' + name], 1e-4, 1e-9, 'grad_' + name)


def test_lightgcl_composition_matches_reference_on_cpu(monkeypatch):
    from sslrec_b200.general_cf import lightgcl as M
    from sslrec_b200 import engine as E
    g, case, adj, dr = _setup('lightgcl')
    hp = g['hp']
    model, _ = H.make_model('lightgcl', case, hp, device='cpu')
    sd = {'user_embeds': case['user_e'], 'item_embeds': case['item_e']}
    sd.update({f'Ws.{i}.W': w for i, w in enumerate(dr['ws'])})
    model.load_state_dict(sd)
    model.ut, model.vt, model.u_mul_s, model.v_mul_s = (torch.from_numpy(g['svd_' + k]) for k in ('ut', 'vt', 'u_mul_s', 'v_mul_s'))
    ladj = O.lightgcl_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    r, c, v = model._ui                                        # the model's own R / sqrt(rowD colD): bit-identical to the reference's
    o, og = np.lexsort((c, r)), np.lexsort((g['lgcl_cols'], g['lgcl_rows']))
    assert np.array_equal(v[o].view(np.uint32), g['lgcl_vals'][og].view(np.uint32))
    monkeypatch.setattr(model, '_bipartite_plan', lambda: None)
    monkeypatch.setattr(E, 'spmm', lambda plan, x, view, layer: torch.sparse.mm(ladj.torch_coo(x.dtype), x))
    monkeypatch.setattr(M, 'cal_bpr_loss', O.bpr_loss_sum)
    monkeypatch.setattr(E, 'dense_logsumexp_mean',
                        lambda a, t, temp, eps=1e-8: torch.log(torch.exp(a @ t.T / temp).sum(1) + eps).mean())
    batch = [torch.from_numpy(case[k]) for k in ('ancs', 'poss', 'negs')]
    loss, parts = model.cal_loss(batch)
    _close(loss.item(), g['loss'], 2e-6, 1e-7, 'loss')
    for k, val in parts.items():
        _close(float(val), g['part_' + k], 2e-6, 1e-9, k)
    loss.backward()
    for name, p in model.named_parameters():
        _close(p.grad, g['grad_' + name], 1e-4, 1e-9, 'grad_' + name)
    assert set(model.state_dict()) == {'user_embeds', 'item_embeds', 'Ws.0.W', 'Ws.1.W'}


def test_exact_order_forward_composition_matches_the_reference_bit_for_bit(monkeypatch):
    """``test.exact_order``: LightGCN._exact_forward composes the layers and the layer sum exactly as the reference does (lightgcn.py:34-42).  With
    the exact-order SpMM replaced by torch's CPU ``sparse.mm`` on the reference's adjacency (which it equals bit for bit,
    tests/test_host_emulation.py) the embeddings are BIT-identical to the oracle's / reference's forward pass, and the eval cache follows
    ``is_training`` like ``forward``'s."""
    from sslrec_b200 import engine as E
    g, case, adj, dr = _setup('lightgcn')
    model, _ = H.make_model('lightgcn', case, g['hp'], device='cpu')
    model.load_state_dict({'user_embeds': case['user_e'], 'item_embeds': case['item_e']})
    coo = adj.torch_coo()
    calls = []

    def spmm_exact(plan, x):
        calls.append(x.shape)
        return torch.sparse.mm(coo, x)
    monkeypatch.setattr(E, 'spmm_exact', spmm_exact)
    monkeypatch.setattr(model, '_plan', lambda adj=None: None)
    u, i = model._exact_forward()
    want = O.lightgcn_embeds(coo, torch.cat([case['user_e'], case['item_e']], 0), g['hp']['layer_num'])
    got = torch.cat([u, i], 0)
    assert np.array_equal(got.numpy().view(np.uint32), want.numpy().view(np.uint32))
    assert len(calls) == g['hp']['layer_num']
    model.is_training = False
    model._exact_forward()
    assert len(calls) == g['hp']['layer_num']                    # served from the cache while evaluating
    model.is_training = True
    model._exact_forward()
    assert len(calls) == 2 * g['hp']['layer_num']                # recomputed after a training step
