"""Multi-GPU plumbing for the NVF path: one process per GPU, torch.distributed.

The path shards over leaf blocks:
* decode: contiguous block ranges per rank (so that concatenating the ranks'
  results in rank order reproduces the single-GPU point order) and one
  variable-length gather of the int32 coordinates to rank 0;
* train (weight loop, NVFPCC.py:149-223): data-parallel over blocks; the only
  exchange is ONE all-reduce of the flattened shared-weight gradient (the
  batch-global `n_pts` is known from per-block counts and the deterministic
  batch schedule, trainer.fit); embedding rows, their Adam state and their
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

    coords [K_r,3] int32, counts [N_r] int32 -> on dst: (coords [sum K_r,3], counts [sum N_r]); elsewhere
    (None, None).  One all-gather of the two sizes (16 bytes per rank), then every other rank SENDS its exact
    tensors and `dst` receives them straight into the slices of the preallocated result (batched point-to-point:
    one NCCL group).  No padding buffers, nothing is delivered to ranks that do not need it, and only `dst`
    reads the sizes on the host (it has to: they are the shape of its result) - the senders never synchronise.
    Payload ~10 MB (vox10) / ~40 MB (vox11) in total."""
    if not is_dist():
        return coords, counts
    rank, ws = world()
    dev = coords.device
    sizes = torch.tensor([coords.shape[0], counts.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = torch.empty((ws, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes) if dev.type == "cuda" else dist.all_gather(
        list(all_sizes.unbind(0)), sizes)
    coords, counts = coords.contiguous(), counts.contiguous()
    if rank != dst:
        ops = []
        if coords.shape[0]:
            ops.append(dist.P2POp(dist.isend, coords, dst))
        if counts.shape[0]:
            ops.append(dist.P2POp(dist.isend, counts, dst))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None, None
    sz = all_sizes.tolist()                                       # dst only: the shape of its result
    k_off = [0]
    n_off = [0]
    for r in range(ws):
        k_off.append(k_off[-1] + sz[r][0])
        n_off.append(n_off[-1] + sz[r][1])
    c = torch.empty((k_off[-1], 3), dtype=coords.dtype, device=dev)
    n = torch.empty((n_off[-1],), dtype=counts.dtype, device=dev)
    c[k_off[dst]:k_off[dst + 1]] = coords
    n[n_off[dst]:n_off[dst + 1]] = counts
    ops = []
    for r in range(ws):
        if r == dst:
            continue
        if sz[r][0]:
            ops.append(dist.P2POp(dist.irecv, c[k_off[r]:k_off[r + 1]], r))
        if sz[r][1]:
            ops.append(dist.P2POp(dist.irecv, n[n_off[r]:n_off[r + 1]], r))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
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
