"""Plugin base class, mirror of models/base_model.py:6-47 (same constructor contract, same
``_mask_predict`` formula), plus the flat embedding table the kernels read without a concat."""
from __future__ import annotations

import torch
from torch import nn

from .config import configs


class BaseModel(nn.Module):
    def __init__(self, data_handler):
        super().__init__()
        self.user_num = configs['data']['user_num']
        self.item_num = configs['data']['item_num']
        self.embedding_size = configs['model']['embedding_size']

    def forward(self):
        pass

    def cal_loss(self, batch_data):
        """-> (0-d loss tensor, dict of loss terms)   (base_model.py:23-33)"""
        pass

    def _mask_predict(self, full_preds, train_mask):
        return full_preds * (1 - train_mask) - 1e8 * train_mask          # base_model.py:35-36

    def full_predict(self, batch_data):
        pass

    # ---- flat table -------------------------------------------------------------------------
    def _alloc_embeddings(self):
        """``user_embeds`` / ``item_embeds`` (the checkpoint contract) as adjacent
        views of ONE [N, d] storage, xavier-uniform per side in the reference's order
        (lightgcn.py:21-22) so the same torch seed gives the same initial weights."""
        table = torch.empty(self.user_num + self.item_num, self.embedding_size)
        nn.init.xavier_uniform_(table[:self.user_num])
        nn.init.xavier_uniform_(table[self.user_num:])
        self.user_embeds = nn.Parameter(table[:self.user_num])
        self.item_embeds = nn.Parameter(table[self.user_num:])

    def _retie(self):
        u, i = self.user_embeds, self.item_embeds
        adjacent = (u.is_contiguous() and i.is_contiguous()
                    and u.untyped_storage().data_ptr() == i.untyped_storage().data_ptr()
                    and i.data_ptr() == u.data_ptr() + u.numel() * u.element_size())
        if not adjacent:
            table = torch.cat([u.data, i.data], 0)
            u.data = table[:self.user_num]
            i.data = table[self.user_num:]

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)        # .to(device) / .cuda() move the two views separately
        if hasattr(self, 'user_embeds') and hasattr(self, 'item_embeds'):
            self._retie()
        return out
