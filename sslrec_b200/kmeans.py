"""KMeansClustering -- mirror of models/aug_utils.py:134-157 (NCL): Lloyd iterations from uniform random
centroids on detached embeddings (ncl.py:26-28), once every ``epoch_period`` epochs.

Each iteration is ``ssl_kmeans_iter``: the reference's materialised [N, K, d] difference tensor, min,
two index_adds and the division become one assignment kernel (distances in the reference's
sum-of-squared-differences form, warp-private shared-memory accumulators) and one ordered reduction
of the per-CTA partials -- deterministic, no floating-point atomics.  The reference always runs 1000
iterations; here the loop stops once an iteration changes no assignment (checked every ``check_every``
iterations): from then on every further Lloyd step reproduces the same centroids.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib


class KMeansClustering:
    def __init__(self, cluster_num, embedding_size, iters: int = 1000, check_every: int = 16):
        self.cluster_num, self.embedding_size, self.iters, self.check_every = cluster_num, embedding_size, iters, check_every
        self.init_centroids = None        # tests inject the reference's t.rand draw (aug_utils.py:147)
        self.last_iters = 0

    def __call__(self, embeds: torch.Tensor):
        if not embeds.is_cuda:
            raise RuntimeError('sslrec_b200.KMeansClustering: CUDA tensors only (no CPU path)')
        dev, K, d = embeds.device, self.cluster_num, self.embedding_size
        embeds = embeds.detach()
        if embeds.dtype != torch.float32 or embeds.stride(-1) != 1:
            embeds = embeds.float().contiguous()
        n = embeds.shape[0]
        if self.init_centroids is not None:
            cents = self.init_centroids.to(dev, torch.float32).contiguous().clone()
        else:
            cents = torch.rand(K, d, device=dev)                                    # aug_utils.py:147
        n_cta, n_warps = C.c_int32(), C.c_int32()
        check(lib.ssl_kmeans_workspace(n, d, K, C.byref(n_cta), C.byref(n_warps)), 'ssl_kmeans_workspace')
        part_sum = torch.empty(n_cta.value, K, d, device=dev)
        part_cnt = torch.empty(n_cta.value, K, device=dev)
        counts = torch.zeros(K, 1, device=dev)
        idxs = torch.full((n,), -1, device=dev, dtype=torch.int64)
        changed = torch.zeros(self.iters + 1, device=dev, dtype=torch.int32)       # one counter per iteration
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream
            it = 0
            while it < self.iters:
                check(lib.ssl_kmeans_iter(embeds.data_ptr(), embeds.stride(0), n, d, K, cents.data_ptr(), idxs.data_ptr(),
                                          part_sum.data_ptr(), part_cnt.data_ptr(), counts.data_ptr(),
                                          changed.data_ptr() + 4 * it, s), 'ssl_kmeans_iter')
                it += 1
                if it % self.check_every == 0 and int(changed[it - 1].item()) == 0:
                    break
        self.last_iters = it
        return cents, idxs, counts
