"""Mirror of models/loss_utils.py for the in-scope functions (same names).  Each accepts either
plain dense tensors with the reference's signature, or row references (:class:`engine.Rows`) plus
the batch indices, in which case the gather is fused into the kernel and gradients go straight to
the propagation's sinks."""
from __future__ import annotations

import torch

from . import engine as E


def cal_bpr_loss(anc_embeds, pos_embeds, neg_embeds=None, *idx):
    """loss_utils.py:7-10.  Dense: (anc [B,d], pos [B,d], neg [B,d]).  Fused: (users Rows, items Rows,
    ancs, poss, negs)."""
    if isinstance(anc_embeds, E.Rows):
        ancs, poss, negs = (neg_embeds,) + idx
        return E.bpr_loss_sum(anc_embeds, pos_embeds, ancs, poss, negs)
    return E.dense_bpr_loss_sum(anc_embeds, pos_embeds, neg_embeds)


def cal_infonce_loss(embeds1, embeds2, all_embeds2, temp=1.0, idx=None, idx2=None):
    """loss_utils.py:30-39.  Dense: ([B,d], [B,d], [N,d], temp).  Fused: (Rows, Rows, Rows, temp, idx)."""
    if isinstance(all_embeds2, E.Rows):
        return E.infonce_loss_sum(embeds1, embeds2, all_embeds2, idx, temp, idx2)
    return E.dense_infonce_loss_sum(embeds1, embeds2, all_embeds2, temp)


def cal_infonce_loss_spec_nodes(embeds1, embeds2, nodes, temp):
    """loss_utils.py:42-51 (dense tensors; embeds1 usually detached as in hccf.py:78)."""
    return E.dense_infonce_spec_nodes_mean(embeds1, embeds2, nodes, temp)


def alignment(x, y, alpha=2, *idx):
    """loss_utils.py:75-79 (alpha = 2 only).  Dense: (x [B,d], y [B,d]).  Fused: (users Rows, items Rows, ancs, poss)
    -- ``alpha`` then holds ``ancs``."""
    if isinstance(x, E.Rows):
        ancs, poss = (alpha,) + idx
        return E.alignment_mean(x, y, ancs, poss)
    if alpha != 2:
        raise NotImplementedError('alignment: only alpha = 2 (the value every caller in the reference uses)')
    return E._DenseAlignFn.apply(x, y)


def uniformity(x, idx=None):
    """loss_utils.py:82-86.  Dense: (x [B,d]).  Fused: (Rows, idx)."""
    if isinstance(x, E.Rows):
        return E.uniformity_log_mean(x, idx)
    return E._DenseUniformFn.apply(x)


def reg_params(model_or_state):
    """loss_utils.py:20-24: sum_W ||W||_2^2 over all parameters.  With a PropState the embedding
    table's part comes from the deterministic reduction kernel and its gradient is fused into the
    last backward layer; any further parameters (HCCF's hyper embeddings) are added with torch."""
    if isinstance(model_or_state, E.PropState):
        return E.table_sumsq(model_or_state)
    model = model_or_state
    st = getattr(model, '_state', None)
    total = 0
    seen = set()
    if st is not None and st.token is not None:
        total = E.table_sumsq(st)
        seen = {id(model.user_embeds), id(model.item_embeds)}
    for w in model.parameters():
        if id(w) not in seen:
            total = total + w.norm(2).square()
    return total
