"""Deterministic synthetic inputs shared by ``oracle/gen_golden.py`` and the tests.
TEST INFRASTRUCTURE ONLY (see oracle/cf_oracle.py header).

Everything is derived from (case name -> sizes, seed) with numpy ``RandomState`` /
``torch.Generator`` streams, so a golden file only has to store the reference's
*outputs*; the inputs are regenerated bit-identically on any box with this image.
"""
from __future__ import annotations

import numpy as np
import torch

# name: (n_user, n_item, n_edge, dim, batch, seed)
CASES = {
    'tiny': (61, 47, 420, 32, 96, 11),
    'small': (700, 500, 9000, 64, 512, 12),
    'mid': (3000, 2200, 42000, 64, 1024, 13),
}


def bipartite_edges(n_user: int, n_item: int, n_edge: int, seed: int):
    """Unique (u, i) pairs with a skewed item popularity; a few users/items are left with
    degree 0 on purpose (the bundled datasets have such rows)."""
    rs = np.random.RandomState(seed)
    dead_u = max(1, n_user // 40)
    dead_i = max(1, n_item // 40)
    pop = 1.0 / np.arange(1, n_item - dead_i + 1) ** 0.8
    pop /= pop.sum()
    got = set()
    while len(got) < n_edge:
        need = n_edge - len(got)
        u = rs.randint(0, n_user - dead_u, size=need * 2)
        i = rs.choice(n_item - dead_i, size=need * 2, p=pop)
        for a, b in zip(u.tolist(), i.tolist()):
            if len(got) >= n_edge:
                break
            got.add((a, b))
    pairs = np.array(sorted(got), dtype=np.int64)
    perm = rs.permutation(len(pairs))          # the bundled COO files are not row-sorted either
    return pairs[perm, 0].copy(), pairs[perm, 1].copy()


def make_case(name: str):
    n_user, n_item, n_edge, dim, batch, seed = CASES[name]
    rows, cols = bipartite_edges(n_user, n_item, n_edge, seed)
    rs = np.random.RandomState(seed + 1000)
    g = torch.Generator().manual_seed(seed + 2000)
    a = float(np.sqrt(6.0 / (n_user + dim)))
    b = float(np.sqrt(6.0 / (n_item + dim)))
    user_e = (torch.rand(n_user, dim, generator=g) * 2 - 1) * a
    item_e = (torch.rand(n_item, dim, generator=g) * 2 - 1) * b
    pick = rs.randint(0, n_edge, size=batch)           # duplicates are likely and intended
    ancs = rows[pick].astype(np.int64)
    poss = cols[pick].astype(np.int64)
    negs = rs.randint(0, n_item, size=batch).astype(np.int64)
    return dict(name=name, n_user=n_user, n_item=n_item, dim=dim, batch=batch, seed=seed,
                rows=rows, cols=cols, user_e=user_e, item_e=item_e, ancs=ancs, poss=poss, negs=negs)


def uniform_stream(seed: int):
    """Generator used for every injected U[0,1) tensor (masks, noise, dropout, k-means init)."""
    return torch.Generator().manual_seed(seed + 3000)


def draw_uniform(gen: torch.Generator, *shape) -> torch.Tensor:
    return torch.rand(*shape, generator=gen)
