"""Replay a golden case through the oracle.  TEST INFRASTRUCTURE ONLY.

``draws(model_key, case)`` regenerates, in the order ``oracle/gen_golden.py`` fed them to the
reference, every injected random tensor; ``oracle_outputs`` runs the oracle on them and returns
the same keys the golden ``.npz`` holds, so tests can compare key by key.  The CUDA parity
tests use ``draws`` too, to drive the kernels with the same masks / noise.
"""
from __future__ import annotations

import ast
import os
from typing import Dict

import numpy as np
import torch

from . import cf_oracle as O
from . import inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def load_golden(model_key: str, case_name: str) -> Dict:
    z = np.load(os.path.join(GOLDEN_DIR, f'{model_key}_{case_name}.npz'), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g['hp'] = ast.literal_eval(str(g.pop('hp_json')))
    return g


def draws(model_key: str, case: Dict, hp: Dict, adj: O.Adj) -> Dict:
    """Injected randomness for one cal_loss call, keyed by meaning.  Edge masks are returned in
    the oracle Adj's (row-major) entry order; the reference consumed them in its own COO order
    (oracle.coo_order_like_reference)."""
    model = model_key.split('_')[0]
    gen = inputs.uniform_stream(case['seed'])
    U, I, D = case['n_user'], case['n_item'], case['dim']
    N, nnz = U + I, adj.nnz
    ref_order = O.coo_order_like_reference(adj)
    keep = hp.get('keep_rate', 1.0)
    L = hp['layer_num']
    d: Dict = {}

    def edge_keep():
        u_ref = inputs.draw_uniform(gen, nnz)
        m_ref = O.keep_mask_from_uniform(u_ref, keep).numpy()
        m = np.zeros(nnz, dtype=bool)
        m[ref_order] = m_ref
        return m

    if model == 'hccf':
        H = hp['hyper_num']
        a = float(np.sqrt(6.0 / (D + H)))
        d['user_w'] = (inputs.draw_uniform(gen, D, H) * 2 - 1) * a
        d['item_w'] = (inputs.draw_uniform(gen, D, H) * 2 - 1) * a
    if model == 'lightgcl':
        a = float(np.sqrt(6.0 / (D + D)))
        d['ws'] = [(inputs.draw_uniform(gen, D, D) * 2 - 1) * a for _ in range(L)]
    if model == 'lightgcn':
        d['edge_keep'] = edge_keep() if keep != 1.0 else None
    elif model == 'simgcl':
        d['uniforms'] = [[inputs.draw_uniform(gen, N, D) for _ in range(L)] for _ in range(2)]
    elif model == 'sgl':
        if hp['augmentation'] == 'edge_drop':
            d['edge_keeps'] = [edge_keep(), edge_keep()]
            d['node_keeps'] = [None, None]
        else:
            d['edge_keeps'] = [None, None]
            d['node_keeps'] = [O.keep_mask_from_uniform(inputs.draw_uniform(gen, N), keep) for _ in range(2)]
    elif model == 'ncl':
        K = hp['cluster_num']
        d['init_user_centroids'] = inputs.draw_uniform(gen, K, D)
        d['init_item_centroids'] = inputs.draw_uniform(gen, K, D)
    elif model == 'hccf':
        d['edge_keeps'], d['hyper_keeps'] = [], []
        for _ in range(L):
            d['edge_keeps'].append(edge_keep())
            ku = (inputs.draw_uniform(gen, U, hp['hyper_num']) + keep).floor()
            ki = (inputs.draw_uniform(gen, I, hp['hyper_num']) + keep).floor()
            d['hyper_keeps'].append((ku, ki))
    return d


def oracle_loss(model_key: str, case: Dict, hp: Dict, adj: O.Adj, dr: Dict, params: Dict, golden: Dict = None):
    """cal_loss of the oracle for this model.  ``params``: dict of leaf tensors."""
    model = model_key.split('_')[0]
    batch = tuple(torch.from_numpy(case[k]) for k in ('ancs', 'poss', 'negs'))
    ue, ie = params['user_embeds'], params['item_embeds']
    if model == 'lightgcn':
        return O.lightgcn_loss(adj, ue, ie, batch, hp['layer_num'], hp['reg_weight'], hp['keep_rate'], dr['edge_keep'])
    if model == 'simgcl':
        u1 = [u.to(ue.dtype) for u in dr['uniforms'][0]]
        u2 = [u.to(ue.dtype) for u in dr['uniforms'][1]]
        return O.simgcl_loss(adj, ue, ie, batch, hp['layer_num'], hp['reg_weight'], hp['cl_weight'],
                             hp['temperature'], hp['eps'], u1, u2)
    if model == 'sgl':
        return O.sgl_loss(adj, ue, ie, batch, hp['layer_num'], hp['reg_weight'], hp['cl_weight'], hp['temperature'],
                          hp['augmentation'], hp['keep_rate'], dr['edge_keeps'], dr['node_keeps'])
    if model == 'ncl':
        dt = ue.dtype
        if golden is not None:       # k-means state injected from the reference run (ncl.py:26-28)
            uc, ic = torch.from_numpy(golden['user_centroids']).to(dt), torch.from_numpy(golden['item_centroids']).to(dt)
            u2c, i2c = torch.from_numpy(golden['user2cluster']), torch.from_numpy(golden['item2cluster'])
        else:
            uc, u2c, _ = O.kmeans(ue.detach(), dr['init_user_centroids'].to(dt))
            ic, i2c, _ = O.kmeans(ie.detach(), dr['init_item_centroids'].to(dt))
        return O.ncl_loss(adj, ue, ie, batch, hp['layer_num'], hp['high_order'], hp['reg_weight'], hp['proto_weight'],
                          hp['struct_weight'], hp['temperature'], uc, u2c, ic, i2c)
    if model == 'directau':
        return O.directau_loss(adj, ue, ie, batch, hp['layer_num'], hp['gamma'])
    if model == 'lightgcl':
        ladj = O.lightgcl_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
        svd = [torch.from_numpy(golden['svd_' + k]) for k in ('ut', 'vt', 'u_mul_s', 'v_mul_s')]      # t.svd_lowrank draw of the reference run
        ws = [params[f'Ws.{i}.W'] for i in range(hp['layer_num'])]
        return O.lightgcl_loss(ladj, ue, ie, ws, batch, hp['layer_num'], hp['reg_weight'], hp['cl_weight'], hp['temp'], *svd)
    if model == 'hccf':
        return O.hccf_loss(adj, ue, ie, params['user_hyper_embeds'], params['item_hyper_embeds'], batch,
                           hp['layer_num'], hp['reg_weight'], hp['cl_weight'], hp['temperature'], hp['keep_rate'],
                           hp['mult'], hp['leaky'], dr['edge_keeps'], dr['hyper_keeps'])
    raise ValueError(model_key)


def clean_embeds(model_key: str, adj: O.Adj, hp: Dict, params: Dict):
    """Evaluation-time embeddings (no augmentation): lightgcn.py:58-60, simgcl.py:57-58, ncl.py:88-90,
    hccf.py:90-92 (keep_rate 1.0 -> no edge drop / dropout)."""
    model = model_key.split('_')[0]
    ue, ie = params['user_embeds'], params['item_embeds']
    a_t = adj.torch_coo(ue.dtype)
    e0 = torch.cat([ue, ie], 0)
    if model == 'hccf':
        e, _, _ = O.hccf_embeds(adj, ue, ie, params['user_hyper_embeds'], params['item_hyper_embeds'], hp['layer_num'],
                                1.0, hp['mult'], hp['leaky'])
        return e
    if model == 'lightgcl':           # lightgcl.py:71-72,127: the cached E of the last training forward (dropout 0 here)
        ladj = O.lightgcl_adjacency(adj.rows[adj.rows < adj.n_user], adj.cols[adj.rows < adj.n_user] - adj.n_user, adj.n_user, adj.n_item)
        zu, zi = torch.zeros(1, adj.n_user, dtype=ue.dtype), torch.zeros(1, adj.n_item, dtype=ue.dtype)     # the SVD branch does not feed E
        eu, ei, _, _ = O.lightgcl_embeds(ladj, ue, ie, hp['layer_num'], zu, zi, zu.T, zi.T)
        return torch.cat([eu, ei], 0)
    e = O.lightgcn_embeds(a_t, e0, hp['layer_num'])
    return e / (hp['layer_num'] + 1) if model == 'directau' else e      # directau.py:33: mean over the layers


def oracle_outputs(model_key: str, case_name: str, dtype=torch.float32, golden: Dict = None) -> Dict:
    golden = golden if golden is not None else load_golden(model_key, case_name)
    hp = golden['hp']
    case = inputs.make_case(case_name)
    adj = O.normalized_adjacency(case['rows'], case['cols'], case['n_user'], case['n_item'])
    dr = draws(model_key, case, hp, adj)
    params = {'user_embeds': case['user_e'].to(dtype).clone().requires_grad_(True),
              'item_embeds': case['item_e'].to(dtype).clone().requires_grad_(True)}
    if 'user_w' in dr:
        params['user_hyper_embeds'] = dr['user_w'].to(dtype).clone().requires_grad_(True)
        params['item_hyper_embeds'] = dr['item_w'].to(dtype).clone().requires_grad_(True)
    for i, w in enumerate(dr.get('ws', [])):
        params[f'Ws.{i}.W'] = w.to(dtype).clone().requires_grad_(True)
    out: Dict = {'adj': adj, 'draws': dr, 'case': case, 'hp': hp}
    loss, parts = oracle_loss(model_key, case, hp, adj, dr, params, golden)
    out['loss'] = loss.detach()
    for k, v in parts.items():
        out['part_' + k] = torch.as_tensor(v).detach()
    loss.backward()
    for k, p in params.items():
        out['grad_' + k] = p.grad.detach().clone()
    with torch.no_grad():
        e = clean_embeds(model_key, adj, hp, params)
        U, I = case['n_user'], case['n_item']
        bt = min(64, U)
        users = torch.arange(bt)
        mask = torch.zeros(bt, I, dtype=torch.int64)
        sel = case['rows'] < bt
        mask[torch.from_numpy(case['rows'][sel]), torch.from_numpy(case['cols'][sel])] = 1
        preds = O.full_predict(e[:U], e[U:], users, mask)
        out['preds'] = preds
        top = torch.topk(preds, k=min(40, I))
        out['topk_idx'], out['topk_val'] = top.indices, top.values
        for k, p in params.items():
            newp, _, _ = O.adam_update(p.detach(), p.grad, torch.zeros_like(p), torch.zeros_like(p), 1, float(golden.get('opt_lr', 1e-3)),
                                       weight_decay=float(golden.get('opt_weight_decay', 0.0)))
            out['new_' + k] = newp
    return out
