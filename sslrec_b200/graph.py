"""Adjacency plan: the reference's ``data_handler.torch_adj`` (an uncoalesced, column-sorted COO
fp32 tensor, data_utils/data_handler_general_cf.py:53-73) converted ONCE to the int32 CSR the
sm_100a propagation kernel walks.  The structure and the values are symmetric (D^-1/2 A D^-1/2
of an undirected bipartite graph), so the same CSR serves the forward SpMM and the transposed
SpMM of the backward pass; only an *injected* edge mask needs the reverse-entry permutation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


class GraphPlan:
    """CSR of (a row block of) the normalised adjacency + the native work lists.

    rows/cols/vals : COO triplets in ANY order (numpy or torch, host or device).
    n             : number of nodes N = |U| + |I| (matrix is N x N).
    row_range     : (r0, r1) global rows owned by this plan (row-sharded multi-GPU); default all.
    ``coo_to_csr`` maps the caller's entry order to CSR positions so masks given in the
    reference's COO order (aug_utils.py:25-30) can be injected.
    """

    def __init__(self, rows, cols, vals, n: int, device: torch.device, row_range=None, need_rev: bool = False, side_split: int = 0):
        rows = _np(rows).astype(np.int64)
        cols = _np(cols).astype(np.int64)
        vals = _np(vals).astype(np.float32)
        self.n = int(n)
        self.device = torch.device(device)
        r0, r1 = (0, self.n) if row_range is None else (int(row_range[0]), int(row_range[1]))
        self.row_offset, self.n_rows = r0, r1 - r0
        if row_range is not None:                              # a shard sorts only its own entries
            sel = (rows >= r0) & (rows < r1)
            rows, cols, vals = rows[sel], cols[sel], vals[sel]
        order = np.lexsort((cols, rows))                       # CSR order: row, then col
        if row_range is None:
            self.coo_to_csr_full = np.empty_like(order)
            self.coo_to_csr_full[order] = np.arange(order.shape[0])
        rows_s, cols_s, vals_s = rows[order], cols[order], vals[order]
        self.entry_lo = 0
        rows_l, cols_l, vals_l = rows_s - r0, cols_s, vals_s
        self.nnz = int(rows_l.shape[0])
        rowptr = np.zeros(self.n_rows + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(np.bincount(rows_l, minlength=self.n_rows))
        if self.nnz >= 2 ** 31 - 1:
            raise ValueError('a plan holds at most 2^31-2 entries; shard the rows')
        self.h_rowptr = np.ascontiguousarray(rowptr.astype(np.int32))
        self.colidx = torch.from_numpy(cols_l.astype(np.int32)).to(self.device)
        self.vals = torch.from_numpy(vals_l).to(self.device)
        self.rev = None
        if need_rev:
            if row_range is not None:
                raise ValueError('injected masks (rev) are a single-GPU debugging aid')
            key = rows_s * self.n + cols_s
            keyt = cols_s * self.n + rows_s
            pos = np.searchsorted(key, keyt)
            if not np.array_equal(key[pos], keyt):
                raise ValueError('adjacency structure is not symmetric')
            self.rev = torch.from_numpy(pos.astype(np.int32)).to(self.device)
        self._handle = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.ssl_plan_create(
                C.byref(self._handle), self.h_rowptr.ctypes.data, self.colidx.data_ptr(), self.vals.data_ptr(),
                self.rev.data_ptr() if self.rev is not None else None,
                self.n_rows, self.n, self.nnz, self.row_offset, int(side_split), stream), 'ssl_plan_create')

    @classmethod
    def from_torch_adj(cls, adj: torch.Tensor, device=None, need_rev: bool = False, side_split: int = 0) -> 'GraphPlan':
        """From the reference's sparse COO tensor (any device).  side_split = |U| orders the work list side by side."""
        idx = adj._indices() if adj.layout == torch.sparse_coo else adj.to_sparse_coo()._indices()
        val = adj._values() if adj.layout == torch.sparse_coo else adj.to_sparse_coo()._values()
        device = device if device is not None else adj.device
        return cls(idx[0].cpu().numpy(), idx[1].cpu().numpy(), val.cpu().numpy(), adj.shape[0], device, need_rev=need_rev, side_split=side_split)

    @property
    def handle(self):
        return self._handle

    def stats(self):
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib.ssl_plan_stats(self._handle, out))
        return dict(items=out[0], split_rows=out[1], segments=out[2], max_row_nnz=out[3])

    def mask_to_csr(self, mask_in_caller_order) -> torch.Tensor:
        """uint8 keep-mask given in the order of the COO triplets passed to the constructor ->
        device uint8 tensor in CSR entry order (what edge_mode 2 reads)."""
        if not hasattr(self, 'coo_to_csr_full'):
            raise ValueError('injected masks are a single-GPU debugging aid (row-sharded plans keep no COO map)')
        m = _np(mask_in_caller_order).astype(np.uint8)
        out = np.empty_like(m)
        out[self.coo_to_csr_full] = m
        return torch.from_numpy(out[self.entry_lo:self.entry_lo + self.nnz].copy()).to(self.device)

    def __del__(self):
        try:
            if getattr(self, '_handle', None) is not None and self._handle.value:
                _lib.lib.ssl_plan_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass


def _np(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)
