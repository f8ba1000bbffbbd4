"""Mirror of the in-scope augmentors of models/aug_utils.py.  In the reference these are nn.Modules
that build a second adjacency (EdgeDrop, :11-31), multiply by a host-built mask (NodeDrop, :33-50) or
add a materialised noise tensor (EmbedPerturb, :118-132).  Here they are *descriptions*
(:class:`engine.ViewSpec`) evaluated inside the propagation kernel; the classes keep the
reference's names and constructor arguments."""
from __future__ import annotations

from .engine import ViewSpec


class EdgeDrop:
    def __init__(self, resize_val: bool = False):
        self.resize_val = resize_val

    def view(self, keep_rate: float, seed: int, per_layer: bool = False, injected=None) -> ViewSpec:
        if keep_rate == 1.0:                                    # aug_utils.py:24
            return ViewSpec(seed=seed)
        scale = (1.0 / keep_rate) if self.resize_val else 1.0   # aug_utils.py:29
        if injected is not None:
            return ViewSpec(edge_mode=2, keep=keep_rate, scale=scale, edge_masks=injected, seed=seed)
        return ViewSpec(edge_mode=1, keep=keep_rate, scale=scale, per_layer_edges=per_layer, seed=seed)


class NodeDrop:
    def view(self, keep_rate: float, seed: int, injected=None) -> ViewSpec:
        if keep_rate == 1.0:                                    # aug_utils.py:46
            return ViewSpec(seed=seed)
        if injected is not None:
            return ViewSpec(node_mode=2, node_keep=keep_rate, node_mask=injected, seed=seed)
        return ViewSpec(node_mode=1, node_keep=keep_rate, seed=seed)


class EmbedPerturb:
    def __init__(self, eps: float):
        self.eps = eps

    def view(self, seed: int, injected=None) -> ViewSpec:
        if injected is not None:
            return ViewSpec(noise_mode=2, noise_u=injected, seed=seed)
        return ViewSpec(noise_mode=1, seed=seed)
