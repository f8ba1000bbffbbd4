"""Row-sharded multi-GPU execution: the adjacency rows and every [N, V, d]
layer output are split into ``world`` contiguous row blocks, one per GPU (one process per GPU,
``torch.distributed`` NCCL over NVLink 5 / NVSwitch).

Exchange steps -- the only collectives on the path:
  * one all-gather of the [N/world, V, d] layer output per propagation layer, forward and backward
    (each rank computes its rows from the full previous layer);
  * per InfoNCE term one all-reduce of the per-anchor partial (row sum, weighted table average)
    [B, d+1], because the table rows (negatives) are sharded and the anchors are replicated;
  * one all-gather of the [N/world, d] gradient block before the (replicated) Adam step.
Blocks are equal-sized (ceil(N / world), the last one padded) so the gathered buffer's first N rows
ARE the full tensor -- no compaction copy.  Everything else (BPR on the replicated batch, the
regulariser, Adam on the replicated table) is rank-local and bit-identical across ranks.
"""
from __future__ import annotations

import torch

from .graph import GraphPlan


class RowShard:
    def __init__(self, dist, rank: int, world: int, n: int):
        self.dist, self.rank, self.world, self.n = dist, rank, world, n
        self.block = (n + world - 1) // world
        self.r0 = min(n, rank * self.block)
        self.r1 = min(n, self.r0 + self.block)

    @property
    def n_local(self) -> int:
        return self.r1 - self.r0

    def make_plan(self, adj: torch.Tensor, device) -> GraphPlan:
        idx, val = adj._indices(), adj._values()
        return GraphPlan(idx[0].cpu().numpy(), idx[1].cpu().numpy(), val.cpu().numpy(), adj.shape[0], device,
                         row_range=(self.r0, self.r1))

    def alloc_rows(self, *tail, device, dtype=torch.float32) -> torch.Tensor:
        """Local output buffer with ``block`` rows (>= n_local) so it can be all-gathered in place."""
        return torch.empty(self.block, *tail, device=device, dtype=dtype)

    def allgather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """local [block, ...] (rows beyond n_local are padding) -> full [N, ...] on every rank."""
        assert local.shape[0] == self.block and local.is_contiguous()
        out = torch.empty((self.world * self.block,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        self.dist.all_gather_into_tensor(out, local)
        return out[:self.n]

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def local_range(self, off: int, n: int):
        """Intersection of the global rows [off, off+n) with this rank's block, as (lo, hi)."""
        lo, hi = max(off, self.r0), min(off + n, self.r1)
        return (lo, max(lo, hi))
