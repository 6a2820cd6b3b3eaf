"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: trajectory sharding, estimate gather,
gradient all-reduce, max-over-ranks timing."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalfilter_b200.distributed import (
    FlatGradients,
    StreamAllReduce,
    allreduce_gradients,
    gather_estimates,
    max_over_ranks,
    shard_batch,
    shard_bounds,
)


def test_shard_bounds_partition():
    for N in (0, 1, 7, 32, 4096, 4097):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(N, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, N, sd = 4, 7, 3  # ragged: 4 + 3 trajectories
        g = torch.Generator().manual_seed(0)
        full = torch.randn(T, N, sd, generator=g)
        obs = {"image": torch.randn(T, N, 2, 2, generator=g), "gripper_pos": torch.randn(T, N, 3, generator=g)}
        mine = shard_batch(full, rank, world)
        lo, hi = shard_bounds(N, rank, world)
        assert mine.shape == (T, hi - lo, sd) and torch.equal(mine, full[:, lo:hi])
        o = shard_batch(obs, rank, world)
        assert o["image"].shape == (T, hi - lo, 2, 2)
        # per-trajectory work on the shard, then gather: must equal the unsharded computation bit for bit
        local = mine * 2.0 + 1.0
        assert torch.equal(gather_estimates(local.contiguous(), N), full * 2.0 + 1.0)
        # gradient all-reduce: mean over ranks of rank-dependent gradients
        lin = torch.nn.Linear(3, 2)
        torch.manual_seed(1)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        n = allreduce_gradients(lin)
        assert n == 8
        for p in lin.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2.0))
        assert max_over_ranks(10.0 + rank, "cpu") == 10.0 + world - 1
        # the training step's exchange (bench.py run_c4, two-graph form): parameter gradients are views into ONE flat
        # buffer; backward accumulates into the views, one all-reduce of the buffer, scale, optimiser step on the views
        torch.manual_seed(7)  # identical replicas on both ranks
        net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
        flat = FlatGradients(net.parameters())
        assert flat.flat.numel() == sum(p.numel() for p in net.parameters()) and flat.intact()
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        xs = torch.randn(world, 6, 5, generator=torch.Generator().manual_seed(3))  # one shard of the batch per rank
        flat.zero()
        net(xs[rank]).pow(2).mean().backward()
        assert flat.intact(), "backward replaced a gradient view"
        local = flat.flat.clone()
        dist.all_reduce(flat.flat, op=dist.ReduceOp.SUM)
        flat.flat.mul_(1.0 / world)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(flat.flat, torch.stack(gathered).mean(dim=0), rtol=1e-6, atol=1e-8)
        before = [p.detach().clone() for p in net.parameters()]
        opt.step()
        for p, b in zip(net.parameters(), before):  # the optimiser saw the averaged gradient through the views
            assert torch.allclose(p.detach(), b - 0.1 * p.grad)
        after = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        both = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(both, after)
        assert torch.equal(both[0], both[1]), "replicas diverged after the exchanged step"
        with pytest.raises(AssertionError):  # the stream-ordered NCCL path refuses a CPU device instead of falling back
            StreamAllReduce(torch.device("cpu"))
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
