"""HCCF -- drop-in for models/general_cf/hccf.py.  The GCN half of every layer is the sm_100a SpMM
with a fresh, rescaled in-kernel edge mask (hccf.py:33,47); the two contrastive terms per layer and
side run through the fused InfoNCE kernels (loss_utils.py:42-51); the hyper-graph half
(E W, H^T E, H . : dense [N_side, d] x [d, hyper_num] products, hccf.py:43-49,100-108) are plain
library GEMMs left to cuBLAS through torch, tied together by torch autograd."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import engine as E
from ..aug_utils import EdgeDrop
from ..base_model import BaseModel
from ..config import configs
from ..loss_utils import cal_bpr_loss, cal_infonce_loss_spec_nodes, reg_params

init = nn.init.xavier_uniform_


class HCCF(BaseModel):
    def __init__(self, data_handler):
        super().__init__(data_handler)
        model_config = configs['model']
        self.adj = data_handler.torch_adj
        self.layer_num = model_config['layer_num']
        self.reg_weight = model_config['reg_weight']
        self.cl_weight = model_config['cl_weight']
        self.hyper_num = model_config['hyper_num']
        self.mult = model_config['mult']
        self.keep_rate = model_config['keep_rate']
        self.temperature = model_config['temperature']

        self._alloc_embeddings()                                                            # hccf.py:27-28
        self.hgnn_layer = HGNNLayer(model_config['leaky'])
        self.user_hyper_embeds = nn.Parameter(init(torch.empty(self.embedding_size, self.hyper_num)))
        self.item_hyper_embeds = nn.Parameter(init(torch.empty(self.embedding_size, self.hyper_num)))
        self.edge_drop = EdgeDrop(resize_val=True)
        self._init_runtime(data_handler)

    def _gcn_layer(self, embeds, view, layer):
        return E.spmm(self._plan(), embeds, view, layer)

    def _dropout(self, x, keep_rate, layer, side):
        if keep_rate == 1.0:
            return x
        if self._inject is not None and 'hyper_keeps' in self._inject:
            keep = self._inject['hyper_keeps'][layer][side]
            return x * keep.to(x.dtype) / keep_rate
        return F.dropout(x, p=1 - keep_rate)

    def forward(self, adj, keep_rate):
        embeds = torch.concat([self.user_embeds, self.item_embeds], dim=0)
        embeds_list = [embeds]
        gcn_embeds_list, hyper_embeds_list = [], []
        uu_hyper = self.user_embeds @ self.user_hyper_embeds * self.mult
        ii_hyper = self.item_embeds @ self.item_hyper_embeds * self.mult
        inj = None if self._inject is None else self._inject.get('edge_masks_per_layer')
        seed = self._seeds.next()
        view = self.edge_drop.view(keep_rate, seed, per_layer=True, injected=inj)             # one mask per layer
        for i in range(self.layer_num):
            tem_embeds = self._gcn_layer(embeds_list[-1], view, i + 1)
            hyper_user_embeds = self.hgnn_layer(self._dropout(uu_hyper, keep_rate, i, 0), embeds_list[-1][:self.user_num])
            hyper_item_embeds = self.hgnn_layer(self._dropout(ii_hyper, keep_rate, i, 1), embeds_list[-1][self.user_num:])
            gcn_embeds_list.append(tem_embeds)
            hyper_embeds_list.append(torch.concat([hyper_user_embeds, hyper_item_embeds], dim=0))
            embeds_list.append(tem_embeds + hyper_embeds_list[-1])
        embeds = sum(embeds_list)
        return embeds, gcn_embeds_list, hyper_embeds_list

    def cal_loss(self, batch_data):
        ancs, poss, negs = batch_data
        embeds, gcn_embeds_list, hyper_embeds_list = self.forward(self.adj, self.keep_rate)
        user_embeds, item_embeds = embeds[:self.user_num], embeds[self.user_num:]
        # -log sigmoid(a.p - a.n).mean() == softplus(a.n - a.p).mean()  (hccf.py:70-74)
        bpr_loss = cal_bpr_loss(user_embeds[ancs], item_embeds[poss], item_embeds[negs]) / ancs.shape[0]
        cl_loss = 0
        ua, up = torch.unique(ancs), torch.unique(poss)
        for i in range(self.layer_num):
            embeds1 = gcn_embeds_list[i].detach()
            embeds2 = hyper_embeds_list[i]
            cl_loss = cl_loss + cal_infonce_loss_spec_nodes(embeds1[:self.user_num], embeds2[:self.user_num], ua, self.temperature) + \
                cal_infonce_loss_spec_nodes(embeds1[self.user_num:], embeds2[self.user_num:], up, self.temperature)
        reg_loss = reg_params(self) * self.reg_weight
        cl_loss = cl_loss * self.cl_weight
        loss = bpr_loss + reg_loss + cl_loss
        losses = {'bpr_loss': bpr_loss, 'reg_loss': reg_loss, 'cl_loss': cl_loss}
        return loss, losses

    def full_predict(self, batch_data):
        embeds, _, _ = self.forward(self.adj, 1.0)
        return self._predict(embeds[:self.user_num], embeds[self.user_num:], batch_data)


class HGNNLayer(nn.Module):
    def __init__(self, leaky):
        super().__init__()
        self.act = nn.LeakyReLU(negative_slope=leaky)

    def forward(self, adj, embeds):
        hids = self.act(adj.T @ embeds)
        return self.act(adj @ hids)
