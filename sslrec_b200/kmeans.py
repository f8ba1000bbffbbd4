"""KMeansClustering -- mirror of models/aug_utils.py:134-157 (NCL).  Lloyd iterations on device.
The assignment step uses |x|^2 - 2 x.c + |c|^2 through a cuBLAS GEMM (a plain library GEMM) instead
of the reference's materialised [N, K, d] difference tensor; the update step is index_add as in the
reference.  It runs once every ``epoch_period`` epochs on detached embeddings (ncl.py:26-28), outside
the per-step hot path; a native kernel is SURVEY.md section 8(f) row 4."""
from __future__ import annotations

import torch


class KMeansClustering:
    def __init__(self, cluster_num, embedding_size, iters: int = 1000):
        self.cluster_num, self.embedding_size, self.iters = cluster_num, embedding_size, iters
        self.init_centroids = None        # tests inject the reference's t.rand draw (aug_utils.py:147)

    def __call__(self, embeds: torch.Tensor):
        dev = embeds.device
        if self.init_centroids is not None:
            cents = self.init_centroids.to(dev).clone()
        else:
            cents = torch.rand(self.cluster_num, self.embedding_size, device=dev)
        ones = torch.ones(embeds.shape[0], 1, device=dev)
        x2 = embeds.square().sum(-1, keepdim=True)
        idxs = None
        prev = None
        for _ in range(self.iters):
            d2 = x2 - 2.0 * embeds @ cents.T + cents.square().sum(-1).unsqueeze(0)
            idxs = d2.argmin(1)
            new = torch.zeros_like(cents).index_add_(0, idxs, embeds)
            cnt = torch.zeros(cents.shape[0], 1, device=dev).index_add_(0, idxs, ones)
            cents = new / (cnt + 1e-6)
            if prev is not None and _ % 16 == 0:
                if torch.equal(prev, idxs):      # converged: further Lloyd steps are fixed points (up to the 1e-6 guard)
                    break
            prev = idxs
        return cents, idxs, cnt
