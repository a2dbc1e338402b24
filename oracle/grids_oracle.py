"""CPU oracle for the ground-truth grid / distance-field builder.

TEST INFRASTRUCTURE ONLY: a restatement of util_get_grids.py:19-46 in numpy.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product package `nvfpcc_b200` never does.

Parity status: PINNED up to the KD-tree.  open3d is not installed here, so
`oracle/gen_golden_grids.py` executes the UNMODIFIED reference script with a stub
`open3d` module whose KDTreeFlann answers `search_knn_vector_3d(p, 1)` by exhaustive
search (the exact answer a KD-tree returns for k=1; ties do not matter because only
the DISTANCE is used, util_get_grids.py:41).  Its outputs are committed as
`tests/golden/grids_small.npz`; `tests/test_grids_oracle.py` checks both functions
below against them.  What stays unpinned is open3d's own FLANN build returning the
true nearest neighbour, which is its documented contract.
"""
from __future__ import annotations

import numpy as np

LEAF = 32


def cube_template(leaf: int = LEAF) -> np.ndarray:
    """util_get_grids.py:19-24: cube_template[i,j,k] = (i,j,k)."""
    g = np.arange(leaf, dtype=np.int64)
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), -1)


def build_grids(points: np.ndarray, origins: np.ndarray, workers: int = 1):
    """util_get_grids.py:26-46 with scipy's exact cKDTree in place of open3d's KDTreeFlann.
    Returns (gt_grid uint8 [N,1,32,32,32], dist float64 [N,1,32,32,32])."""
    from scipy.spatial import cKDTree

    refs = np.asarray(points, dtype=np.float64)
    tree = cKDTree(refs)
    tmpl = cube_template().reshape(-1, 3)
    n = origins.shape[0]
    dist = np.empty((n, LEAF ** 3), dtype=np.float64)
    for s in range(0, n, 64):
        o = np.asarray(origins[s:s + 64], dtype=np.int64)
        q = (o[:, None, :] + tmpl[None, :, :]).reshape(-1, 3)                 # :26-30
        _, idx = tree.query(q.astype(np.float64), k=1, workers=workers)       # :37-39
        d = np.sqrt(np.sum(np.square(refs[idx] - q), -1))                     # :41
        dist[s:s + 64] = d.reshape(-1, LEAF ** 3)
    dist = dist.reshape(n, 1, LEAF, LEAF, LEAF)                               # :42
    gt = (dist == 0).astype(np.uint8)                                         # :43
    return gt, dist


def build_grids_bruteforce(points: np.ndarray, origins: np.ndarray):
    """Same result by exhaustive integer search (small inputs only): the definition
    the KD-tree implements."""
    pts = np.asarray(points, dtype=np.int64)
    tmpl = cube_template().reshape(-1, 3)
    n = origins.shape[0]
    d2 = np.empty((n, LEAF ** 3), dtype=np.int64)
    for b in range(n):
        q = np.asarray(origins[b], dtype=np.int64)[None, :] + tmpl
        best = np.full(q.shape[0], np.iinfo(np.int64).max)
        for s in range(0, pts.shape[0], 256):
            diff = q[:, None, :] - pts[None, s:s + 256, :]
            best = np.minimum(best, (diff * diff).sum(-1).min(1))
        d2[b] = best
    dist = np.sqrt(d2.astype(np.float64)).reshape(n, 1, LEAF, LEAF, LEAF)
    return (dist == 0).astype(np.uint8), dist
