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

``BatchShard`` is the other way to use N GPUs: the batches are the sharded unit.  Every rank runs the
whole step on its OWN batch of B samples and the only exchange is one all-reduce (average) of the
parameter gradients before the Adam step.  Every loss term of the path is a batch sum divided by the
batch size (lightgcn.py:52, simgcl.py:48-50, sgl.py:56-60, ncl.py:58,68,82) or does not depend on
the batch (reg_params), and the in-kernel augmentation draws are keyed by the shared seed, so the
averaged gradient IS the gradient of one reference step at ``batch_size = world * B``: the optimiser
trajectory is that of the reference with the larger batch, and the parameters stay identical on all
ranks.  (HCCF's contrastive term averages over the batch's UNIQUE nodes, hccf.py:80-81, so there the
average of per-rank terms is the usual data-parallel approximation, not an identity.)
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


class BatchShard:
    """Data-parallel training over batches (module docstring).  ``average_gradients`` is called between
    ``loss.backward()`` and ``optimizer.step()`` (trainer.py:67-68)."""

    def __init__(self, dist, rank: int, world: int):
        self.dist, self.rank, self.world = dist, rank, world
        self._avg = dist.get_backend() == 'nccl'          # gloo has no ReduceOp.AVG

    @staticmethod
    def coalesce(grads):
        """Merge gradients that are adjacent contiguous views of one storage (user_embeds.grad / item_embeds.grad
        are the two halves of the flat [N, d] sink) into single 1-D views: one collective instead of several."""
        flat, out = [], []                       # out: (position of the group's first member in ``grads``, buffer)
        for pos, g in enumerate(grads):
            if g.is_contiguous() and g.dtype == torch.float32:
                flat.append((pos, g))
            else:
                out.append((pos, g))
        flat.sort(key=lambda e: (e[1].untyped_storage().data_ptr(), e[1].data_ptr()))
        k = 0
        while k < len(flat):
            first, n, pos = flat[k][1], flat[k][1].numel(), flat[k][0]
            j = k + 1
            while (j < len(flat) and flat[j][1].untyped_storage().data_ptr() == first.untyped_storage().data_ptr()
                   and flat[j][1].data_ptr() == first.data_ptr() + 4 * n):
                n += flat[j][1].numel()
                pos = min(pos, flat[j][0])
                j += 1
            out.append((pos, first if j == k + 1 else torch.as_strided(first, (n,), (1,))))
            k = j
        # addresses differ between processes: the collectives must be issued in parameter order on every rank
        return [b for _, b in sorted(out, key=lambda e: e[0])]

    def average_gradients(self, params) -> None:
        grads = [p.grad for p in params if p.grad is not None]
        for g in self.coalesce(grads):
            if self._avg:
                self.dist.all_reduce(g, op=self.dist.ReduceOp.AVG)
            else:
                self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM)
                g.mul_(1.0 / self.world)

    def shard_loader(self, dataset, batch_size: int, seed: int = 0):
        """The train DataLoader of data_handler_general_cf.py:95 with this rank's 1/world share of every epoch's
        shuffle (call ``loader.sampler.set_epoch(e)`` per epoch, as Trainer.train_epoch does)."""
        from torch.utils import data
        sampler = data.distributed.DistributedSampler(dataset, num_replicas=self.world, rank=self.rank, shuffle=True, seed=seed)
        return data.DataLoader(dataset, batch_size=batch_size, sampler=sampler, num_workers=0)
