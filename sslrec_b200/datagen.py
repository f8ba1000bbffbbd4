"""Synthetic bipartite interaction graphs with the shape statistics of the reference's datasets
(SURVEY.md section 8d): user degree ~ lognormal(sigma = 1) scaled to the requested mean and clipped
to [1, 1e4], items drawn by Zipf(alpha) popularity over a random permutation of the item ids,
duplicates removed, exactly ``n_edge`` unique (user, item) pairs.  numpy PCG64, fixed seed."""
from __future__ import annotations

import numpy as np

# |U|, |I|, E of the reference's bundled train matrices and BASELINE.json config 4
SHAPES = {
    'gowalla': (25557, 19747, 294983),
    'yelp': (42712, 26822, 182357),
    'amazon': (76469, 83761, 966680),
    'synthetic-xl': (10_000_000, 2_000_000, 300_000_000),
    'synthetic-xl-8th': (1_250_000, 250_000, 37_500_000),      # one GPU's eighth of config 4 (same degree statistics)
}
# item-popularity exponent: 0.5 reproduces the bundled datasets' head (max item degree ~1e3 at
# amazon's size; the real matrices have 841 / 309 / 1018); 1.0 is BASELINE.json config 4's generator
ZIPF = {'gowalla': 0.5, 'yelp': 0.5, 'amazon': 0.5, 'synthetic-xl': 1.0, 'synthetic-xl-8th': 1.0}


def named_graph(name: str, seed: int = 2023):
    n_user, n_item, n_edge = SHAPES[name]
    rows, cols = bipartite_graph(n_user, n_item, n_edge, seed, ZIPF[name])
    return rows, cols, n_user, n_item


def _merge_unique(keys: np.ndarray, new: np.ndarray) -> np.ndarray:
    """sorted-unique union of a sorted-unique array with arbitrary new values (one sort of ``new`` only)."""
    new.sort()
    if new.size:
        new = new[np.concatenate([[True], new[1:] != new[:-1]])]
    if keys.size == 0:
        return new
    pos = np.searchsorted(keys, new)
    fresh = (pos == keys.size) | (keys[np.minimum(pos, keys.size - 1)] != new)
    return np.insert(keys, pos[fresh], new[fresh])


def bipartite_graph(n_user: int, n_item: int, n_edge: int, seed: int = 2023, zipf_alpha: float = 1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    deg = rng.lognormal(mean=0.0, sigma=1.0, size=n_user)
    deg = np.clip(deg * (n_edge / deg.sum()), 1, 1e4)
    deg = np.maximum(1, np.round(deg * (n_edge / deg.sum()))).astype(np.int64)
    pop = 1.0 / np.arange(1, n_item + 1, dtype=np.float64) ** zipf_alpha
    cdf = np.cumsum(pop)
    cdf /= cdf[-1]
    perm = rng.permutation(n_item)
    keys = np.empty(0, dtype=np.int64)
    want = deg.copy()
    for it in range(12):
        users = np.repeat(np.arange(n_user, dtype=np.int64), want)
        if it < 2:          # popularity-driven draws; later passes fill the collision losses uniformly
            items = perm[np.searchsorted(cdf, rng.random(users.shape[0]), side='right').clip(0, n_item - 1)]
        else:
            items = rng.integers(0, n_item, size=users.shape[0], dtype=np.int64)
        keys = _merge_unique(keys, users * n_item + items)
        if keys.shape[0] >= n_edge:
            break
        missing = n_edge - keys.shape[0]
        have = np.bincount(keys // n_item, minlength=n_user)
        want = np.maximum(deg - have, 0)
        short = missing - int(want.sum())
        if short > 0:       # rounding of the degree targets: spread the remainder over random users
            want = want + np.bincount(rng.integers(0, n_user, size=int(short * 1.05) + 1), minlength=n_user)
    if keys.shape[0] > n_edge:
        keys = np.sort(rng.choice(keys, size=n_edge, replace=False))
    if keys.shape[0] != n_edge:
        raise RuntimeError('could not place %d unique edges' % n_edge)
    return (keys // n_item).astype(np.int64), (keys % n_item).astype(np.int64)
