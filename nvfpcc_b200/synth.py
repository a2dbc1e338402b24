"""Synthetic inputs for tests and benchmarks.

The reference's data (longdress_vox10_1300.ply, SEED3.npy) is not available
offline, so benchmarks and tests use a voxelised sphere shell cut into 32^3
leaf blocks, laid out exactly as the reference's preprocessing would emit it:

* leaf origins in the octree's DFS order (get_octree.cpp:368-371, 788-795:
  child index = x + 2y + 4z, i.e. Morton order with z most significant),
* `gt_grid[n,0,i,j,k]` = occupancy of voxel origin[n] + (i,j,k) and
  `dist[n,0,i,j,k]` = distance to the nearest cloud point
  (util_get_grids.py:19-46).
"""
from __future__ import annotations

import numpy as np

LEAF = 32


def synthetic_seed(n: int = 300_000) -> np.ndarray:
    """Stand-in for the reference's SEED3.npy (utils/network.py:20)."""
    return np.random.default_rng(0).random(n)


def sphere_shell_points(resolution: int = 1024, radius: float | None = None) -> np.ndarray:
    """Integer voxels p with | ||p - c|| - R | < 0.5, c = resolution/2.
    resolution 1024, R 260 -> 849 338 points (vox10); 2048, R 520 -> vox11."""
    c = resolution // 2
    R = float(radius if radius is not None else 260.0 * resolution / 1024.0)
    lo = int(np.floor(c - R - 1))
    hi = int(np.ceil(c + R + 1))
    ax = np.arange(lo, hi + 1, dtype=np.int64)
    X, Y = np.meshgrid(ax, ax, indexing="ij")
    rxy2 = (X - c) ** 2 + (Y - c) ** 2
    r_in2 = (R - 0.5) ** 2
    r_out2 = (R + 0.5) ** 2
    pts = []
    # |dz| ranges from ceil(sqrt(r_in2 - rxy2))-ish to floor(sqrt(r_out2 - rxy2)); enumerate a short band
    dz_hi = np.floor(np.sqrt(np.maximum(r_out2 - rxy2, -1.0).clip(min=0))).astype(np.int64)
    dz_hi[r_out2 - rxy2 <= 0] = -1
    dz_lo = np.ceil(np.sqrt(np.maximum(r_in2 - rxy2, 0.0))).astype(np.int64)
    band = int((dz_hi - dz_lo).max()) + 1
    for k in range(band + 1):
        dz = dz_lo + k
        d2 = rxy2 + dz * dz
        ok = (dz <= dz_hi) & (d2 > r_in2) & (d2 < r_out2)
        if not ok.any():
            continue
        x, y, z = X[ok], Y[ok], dz[ok]
        pts.append(np.stack([x, y, c + z], 1))
        nz = z != 0
        pts.append(np.stack([x[nz], y[nz], c - z[nz]], 1))
    p = np.concatenate(pts, 0)
    p = np.unique(p, axis=0)
    return p.astype(np.int32)


def _morton_key(cells: np.ndarray, bits: int) -> np.ndarray:
    """DFS order of get_octree.cpp: at each level child = x + 2y + 4z, so the
    sort key interleaves bits with z most significant."""
    key = np.zeros(cells.shape[0], dtype=np.int64)
    x, y, z = (cells[:, i].astype(np.int64) for i in range(3))
    for b in range(bits - 1, -1, -1):
        key = (key << 3) | (((z >> b) & 1) << 2) | (((y >> b) & 1) << 1) | ((x >> b) & 1)
    return key


def leaf_origins(points: np.ndarray, leaf: int = LEAF) -> np.ndarray:
    """Origins (N,3) int32 of occupied leaf cubes, in octree DFS order."""
    cells = np.unique(points // leaf, axis=0)
    bits = max(1, int(np.ceil(np.log2(cells.max() + 1))))
    order = np.argsort(_morton_key(cells, bits), kind="stable")
    return (cells[order] * leaf).astype(np.int32)


def gt_and_dist(points: np.ndarray, origins: np.ndarray, leaf: int = LEAF, workers: int = -1):
    """util_get_grids.py:19-46 with scipy's cKDTree in place of open3d's
    KDTreeFlann: returns (gt_grid uint8 [N,1,L,L,L], dist float32 [N,1,L,L,L])."""
    from scipy.spatial import cKDTree

    tree = cKDTree(points.astype(np.float64))
    g = np.arange(leaf)
    cube = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    n = origins.shape[0]
    dist = np.empty((n, leaf ** 3), dtype=np.float32)
    chunk = 64
    for s in range(0, n, chunk):
        o = origins[s:s + chunk].astype(np.float64)
        q = (o[:, None, :] + cube[None, :, :]).reshape(-1, 3)
        d, _ = tree.query(q, k=1, workers=workers)
        dist[s:s + chunk] = d.reshape(-1, leaf ** 3).astype(np.float32)
    gt = (dist == 0).astype(np.uint8)
    shape = (n, 1, leaf, leaf, leaf)
    return gt.reshape(shape), dist.reshape(shape)


def random_latents(n: int, ch: int, seed: int = 0) -> np.ndarray:
    """Decode-side latents: round(N(0,3)) clipped to the arithmetic coder's
    symbol range [-512, 511] (NVFPCC.py:446-470), float32 (n,ch,2,2,2)."""
    rng = np.random.default_rng(seed)
    z = np.rint(rng.normal(0.0, 3.0, size=(n, ch, 2, 2, 2)))
    return np.clip(z, -512, 511).astype(np.float32)


def random_kernel_deltas(state: dict, seed: int = 1, sigma: float = 0.05, quantize: bool = True) -> dict:
    """`kernel` deltas for the '4-bit quantised' configuration:
    round(N(0,sigma)*16)/16 on the 7 quantised reconstructor kernels
    (manipulate_weights.py:36-51)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    out = dict(state)
    for name in ("up0", "conv0", "up1", "conv1", "up2", "conv2", "conv2_cls"):
        k = "reconstructor.%s.kernel" % name
        d = torch.randn(state[k].shape, generator=g) * sigma
        out[k] = torch.round(d * 16) / 16 if quantize else d
    return out
