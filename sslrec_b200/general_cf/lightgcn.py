"""LightGCN -- drop-in for models/general_cf/lightgcn.py (class name, ctor, forward, cal_loss,
full_predict, ``is_training`` / ``final_embeds`` cache semantics are the reference's)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import engine as E
from .._lib import check, lib
from ..aug_utils import EdgeDrop
from ..base_model import BaseModel
from ..config import configs
from ..graph import GraphPlan
from ..loss_utils import cal_bpr_loss, reg_params


class LightGCN(BaseModel):
    def __init__(self, data_handler):
        super().__init__(data_handler)
        self.adj = data_handler.torch_adj
        self._trn_mat = getattr(data_handler, 'trn_mat', None)

        self.layer_num = configs['model']['layer_num']
        self.reg_weight = configs['model']['reg_weight']
        self.keep_rate = configs['model']['keep_rate']

        self._alloc_embeddings()                      # user_embeds, item_embeds  (lightgcn.py:21-22)

        self.edge_dropper = EdgeDrop()
        self.is_training = True
        self.final_embeds = None

        self._seeds = E.SeedStream(configs.get('train', {}).get('seed', 2023))
        self._plans = {}
        self._state = None
        self._inject = None        # tests: dict of injected masks / noise (see tests/)
        self.comm = None           # row-sharded multi-GPU communicator (parallel.RowShard), optional

    # ---- adjacency plan (built once per adjacency tensor and device) ---------------------------
    def _plan(self, adj=None) -> GraphPlan:
        adj = self.adj if adj is None else adj
        dev = self.user_embeds.device
        key = (id(adj), str(dev))
        if key not in self._plans:
            if dev.type != 'cuda':
                raise RuntimeError('sslrec_b200 models run on CUDA only (move the model with .to("cuda"))')
            need_rev = self._inject is not None
            if self.comm is not None:
                self._plans[key] = self.comm.make_plan(adj, dev)
            else:
                self._plans[key] = GraphPlan.from_torch_adj(adj, dev, need_rev=need_rev)
        return self._plans[key]

    def _table(self) -> torch.Tensor:
        return E.flat_table(self.user_embeds, self.item_embeds)

    def _propagate(self, views, n_layers=None, sum_layers=None, keep_layers=(), noise_eps=0.0, adj=None) -> E.PropState:
        """All augmented views, all layers and the layer sum: replaces the loop of t.spmm calls
        (lightgcn.py:28-29,38-41)."""
        prop = E.Propagation(self._plan(adj), views, self.layer_num if n_layers is None else n_layers,
                             sum_layers, keep_layers, noise_eps, comm=self.comm)
        st = E.propagate(prop, self.user_embeds, self.item_embeds, self._table())
        self._state = st
        return st

    def _edge_view(self, keep_rate, slot=0):
        inj = None if self._inject is None else self._inject.get('edge_masks', [None] * 4)[slot]
        return self.edge_dropper.view(keep_rate, self._seeds.next(), injected=inj)

    def forward(self, adj, keep_rate):
        if not self.is_training and self.final_embeds is not None:
            return self.final_embeds[:self.user_num], self.final_embeds[self.user_num:]
        view = self._edge_view(keep_rate) if self.is_training else E.ViewSpec()
        st = self._propagate([view], adj=adj)
        embeds = st.E.view(st.n, st.dim)
        self.final_embeds = embeds
        return embeds[:self.user_num], embeds[self.user_num:]

    def cal_loss(self, batch_data):
        self.is_training = True
        ancs, poss, negs = batch_data
        st = self._propagate([self._edge_view(self.keep_rate)])
        self.final_embeds = st.E.view(st.n, st.dim)
        bpr_loss = cal_bpr_loss(st.users(0), st.items(0), ancs, poss, negs) / ancs.shape[0]
        reg_loss = self.reg_weight * reg_params(self)
        loss = bpr_loss + reg_loss
        losses = {'bpr_loss': bpr_loss, 'reg_loss': reg_loss}
        return loss, losses

    def _predict(self, user_embeds, item_embeds, batch_data):
        """E_u[users] E_i^T with the training positives masked to -1e8 (lightgcn.py:61-65,
        base_model.py:35-36), one kernel, no [Bt, I] temporaries besides the result."""
        pck_users, train_mask = batch_data
        pck_users = pck_users.long().contiguous()
        n_b = pck_users.shape[0]
        preds = torch.empty(n_b, self.item_num, device=user_embeds.device, dtype=torch.float32)
        mask = None
        if train_mask is not None:
            mask = train_mask.long().contiguous()
        with torch.cuda.device(preds.device):
            check(lib.ssl_predict_mask(user_embeds.data_ptr(), user_embeds.stride(0), item_embeds.data_ptr(), item_embeds.stride(0),
                                       pck_users.data_ptr(), n_b, self.item_num, self.embedding_size,
                                       None if mask is None else mask.data_ptr(), None, None, preds.data_ptr(),
                                       torch.cuda.current_stream(preds.device).cuda_stream), 'ssl_predict_mask')
        return preds

    def full_predict(self, batch_data):
        user_embeds, item_embeds = self.forward(self.adj, 1.0)
        self.is_training = False
        return self._predict(user_embeds, item_embeds, batch_data)
