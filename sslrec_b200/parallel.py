"""Row-sharded multi-GPU execution: the adjacency rows and every [N, V, d]
layer output are split into ``world`` contiguous row blocks, one per GPU (one process per GPU,
``torch.distributed`` NCCL over NVLink 5 / NVSwitch).

Exchange steps -- the only collectives on the path:
  * one all-gather of the [N/world, V, d] layer output per propagation layer, forward and backward
    (each rank computes its rows from the full previous layer);
  * per InfoNCE term one all-reduce of the per-anchor partial (row sum, weighted table average)
    [B, d+1], because the table rows (negatives) are sharded and the anchors are replicated, and in the
    backward one all-gather of the [N_side/world, d] dense table-gradient blocks;
  * one all-gather of the [N/world, d] gradient block before the (replicated) Adam step.
The propagation all-gathers only happen when the propagation itself is sharded (``shard_propagation``:
automatic by table size); small graphs replicate the sub-millisecond SpMM and shard only the loss.
Blocks are equal-sized (ceil(N / world), the last one padded) so the gathered buffer's first N rows
ARE the full tensor -- no compaction copy.  Everything else (BPR on the replicated batch, the
regulariser, Adam on the replicated table) is rank-local and bit-identical across ranks.
"""
from __future__ import annotations

import torch

from .graph import GraphPlan


class RowShard:
    def __init__(self, dist, rank: int, world: int, n: int, shard_propagation='auto', dim: int = 64, views: int = 3):
        self.dist, self.rank, self.world, self.n = dist, rank, world, n
        self.block = (n + world - 1) // world
        self.r0 = min(n, rank * self.block)
        self.r1 = min(n, self.r0 + self.block)
        # Row-sharding the propagation costs one all-gather of the whole [N, V, d] layer per layer and
        # direction; it pays when the SpMM is long (HBM-bound tables far beyond L2, BASELINE.json config 4),
        # not when a layer takes ~0.2 ms (the bundled datasets).  The contraction of the contrastive loss
        # is sharded over the table rows in either mode.
        if shard_propagation == 'auto':
            shard_propagation = n * views * dim * 4 >= (1 << 30)
        self.shard_propagation = bool(shard_propagation)

    @property
    def n_local(self) -> int:
        return self.r1 - self.r0

    def make_plan(self, adj: torch.Tensor, device, side_split: int = 0) -> GraphPlan:
        idx, val = adj._indices(), adj._values()
        return GraphPlan(idx[0].cpu().numpy(), idx[1].cpu().numpy(), val.cpu().numpy(), adj.shape[0], device,
                         row_range=(self.r0, self.r1), side_split=side_split)

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

    def side_block(self, n: int) -> int:
        return (n + self.world - 1) // self.world

    def side_range(self, off: int, n: int):
        """This rank's share of the table rows [off, off+n) of one side (equal blocks, last padded)."""
        blk = self.side_block(n)
        lo = min(off + n, off + self.rank * blk)
        return lo, min(off + n, lo + blk)

    def allgather_side(self, local: torch.Tensor, n: int) -> torch.Tensor:
        """local [side_block(n), d] -> [n, d] (rows of all ranks' side blocks in order)."""
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        self.dist.all_gather_into_tensor(out, local.contiguous())
        return out[:n]
