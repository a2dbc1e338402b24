"""Multi-GPU plumbing for the NVF path: one process per GPU, torch.distributed.

The path shards over leaf blocks:
* decode: contiguous block ranges per rank (so that concatenating the ranks'
  results in rank order reproduces the single-GPU point order) and one
  variable-length gather of the int32 coordinates;
* train (weight loop, NVFPCC.py:149-223): data-parallel over blocks; the only
  exchange is ONE all-reduce of the flattened shared-weight gradient plus the
  batch-global `n_pts` scalar; embedding rows, their Adam state and their
  gt/dist shards stay rank-local (embedding loop, NVFPCC.py:225-251: no exchange).
Works with backend "nccl" on GPUs and "gloo" on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def block_range(n_blocks: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced range [lo, hi): the first n % world ranks get one extra block."""
    base, rem = divmod(int(n_blocks), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_points(coords: torch.Tensor, counts: torch.Tensor, dst: int = 0):
    """Variable-length gather of per-rank results to `dst` in rank order.

    coords [K_r,3] int32, counts [N_r] int32 -> on dst: (coords [sum K_r,3], counts [sum N_r]);
    elsewhere (None, None).  Uses all_gather of the sizes + padded all_gather (a single
    NCCL collective each; the payload is ~10-40 MB in total)."""
    if not is_dist():
        return coords, counts
    rank, ws = world()
    dev = coords.device
    sizes = torch.tensor([coords.shape[0], counts.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(ws)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu()
    kmax, nmax = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    cpad = torch.zeros((kmax, 3), dtype=coords.dtype, device=dev)
    cpad[:coords.shape[0]] = coords
    npad = torch.zeros((nmax,), dtype=counts.dtype, device=dev)
    npad[:counts.shape[0]] = counts
    cg = [torch.empty_like(cpad) for _ in range(ws)]
    ng = [torch.empty_like(npad) for _ in range(ws)]
    dist.all_gather(cg, cpad)
    dist.all_gather(ng, npad)
    if rank != dst:
        return None, None
    c = torch.cat([cg[r][:int(all_sizes[r, 0])] for r in range(ws)], 0)
    n = torch.cat([ng[r][:int(all_sizes[r, 1])] for r in range(ws)], 0)
    return c, n


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_grads_(params: Iterable[torch.nn.Parameter]) -> int:
    """Sum the shared-weight gradients over ranks with ONE collective on a flat buffer
    (52 219 / 200 915 floats).  Returns the number of floats reduced."""
    ps: List[torch.nn.Parameter] = [p for p in params if p.grad is not None]
    if not ps or not is_dist():
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
    return o
