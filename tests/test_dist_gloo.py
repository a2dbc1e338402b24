"""Multi-process host logic of the block-sharded path on CPU tensors:
world_size 2, gloo backend, rendezvous on 127.0.0.1.  The data path itself has no
collective: decode = contiguous block ranges + one coordinate gather; train = one
all-reduce of the shared-weight gradient + the batch-global n_pts scalar."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nvfpcc_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def test_block_range_partitions_exactly():
    for n in (0, 1, 7, 1247, 5000):
        for w in (1, 2, 3, 4, 8):
            r = [D.block_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _points_of(block):
    """deterministic fake per-block result: `block % 5` points (block, k, 7)."""
    k = block % 5
    return np.stack([np.full(k, block), np.arange(k), np.full(k, 7)], 1).astype(np.int32).reshape(k, 3)


def _gather_case(rank, world):
    n = 23
    lo, hi = D.block_range(n, rank, world)
    pts = [_points_of(b) for b in range(lo, hi)]
    coords = torch.from_numpy(np.concatenate(pts, 0).reshape(-1, 3))
    counts = torch.tensor([p.shape[0] for p in pts], dtype=torch.int32)
    c, k = D.gather_points(coords, counts, dst=0)
    if rank != 0:
        return c is None and k is None
    want = np.concatenate([_points_of(b) for b in range(n)], 0)
    return bool(np.array_equal(c.numpy(), want) and k.tolist() == [b % 5 for b in range(n)])


def test_gather_points_reproduces_single_rank_order():
    """concatenating the ranks' block ranges in rank order == the single-process point order
    (NVFPCC.py:631-638: block order, then in-block order)."""
    assert all(_run(_gather_case).values())


def _gather_ragged_case(rank, world):
    """rank 1 holds nothing at all: the gather must not post zero-sized transfers or hang."""
    if rank == 0:
        coords = torch.arange(12, dtype=torch.int32).reshape(4, 3)
        counts = torch.tensor([1, 3], dtype=torch.int32)
    else:
        coords = torch.zeros((0, 3), dtype=torch.int32)
        counts = torch.zeros((0,), dtype=torch.int32)
    c, k = D.gather_points(coords, counts, dst=0)
    if rank != 0:
        return c is None
    return bool(torch.equal(c, coords) and torch.equal(k, counts))


def test_gather_points_with_an_empty_rank():
    assert all(_run(_gather_ragged_case).values())


def test_epoch_schedule_is_the_dataloader_with_a_short_last_batch():
    """trainer.epoch_schedule: single process = DataLoader(batch_size, drop_last=False) over the stride-2113
    dataset order (utils/dataloader.py:163-167, NVFPCC.py:109-111); sharded = every leaf exactly once per epoch,
    equal step counts, short / empty batches only at the end."""
    from nvfpcc_b200 import trainer
    n, B = 1247, 16
    one = trainer.epoch_schedule(n, 0, 1, B)
    flat = [i for b in one for i in b]
    assert flat == [(i * 2113) % n for i in range(n)] and sorted(flat) == list(range(n))
    assert [len(b) for b in one] == [16] * 77 + [15]
    for world in (2, 3, 8):
        per = B // world if B % world == 0 else 4
        scheds = [trainer.epoch_schedule(n, r, world, per) for r in range(world)]
        assert len({len(s) for s in scheds}) == 1
        seen = []
        for r, sch in enumerate(scheds):
            lo, hi = D.block_range(n, r, world)
            sizes = [len(b) for b in sch]
            assert all(a >= b for a, b in zip(sizes, sizes[1:])) and sizes[0] == per
            seen += [lo + i for b in sch for i in b]
            assert all(0 <= i < hi - lo for b in sch for i in b)
        assert sorted(seen) == list(range(n))


def _grad_case(rank, world):
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    x = torch.arange(10, dtype=torch.float32).reshape(2, 5) + rank
    lin(x).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    n = D.allreduce_grads_(lin.parameters())
    # reference: sum of both ranks' gradients computed locally
    tot = [torch.zeros_like(g) for g in local]
    for r in range(world):
        lin.zero_grad()
        xr = torch.arange(10, dtype=torch.float32).reshape(2, 5) + r
        lin(xr).sum().backward()
        for t, p in zip(tot, lin.parameters()):
            t += p.grad
    n_pts = D.allreduce_sum_(torch.tensor(float(10 + rank)))
    return n == sum(t.numel() for t in tot) and float(n_pts) == 21.0, [t.tolist() for t in tot], \
        [g.tolist() for g in local]


def test_allreduce_grads_sums_shared_weights():
    out = _run(_grad_case)
    assert all(v[0] for v in out.values())


def _allreduce_result(rank, world):
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 2)
    x = torch.ones(3, 4) * (rank + 1)
    lin(x).sum().backward()
    D.allreduce_grads_(lin.parameters())
    return [p.grad.tolist() for p in lin.parameters()]


def test_allreduce_grads_identical_on_all_ranks():
    out = _run(_allreduce_result)
    assert out[0] == out[1]
    # weight grad = sum over ranks of sum_batch x = 3*(1+2) per input, for each of the 2 outputs
    assert np.allclose(np.array(out[0][0]), 9.0) and np.allclose(np.array(out[0][1]), 6.0)
