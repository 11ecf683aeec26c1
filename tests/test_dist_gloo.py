"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, max-over-ranks timing and
the flat-bucket gradient averaging used by the data-parallel training step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from medmamba_b200 import dist as mdist
    r, w, _ = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # shards partition the batch
    lo, hi = mdist.shard_range(11, rank, world)
    sizes = [None] * world
    dist.all_gather_object(sizes, (lo, hi))
    # max over ranks
    mx = mdist.max_over_ranks([10.0 + rank, 5.0 - rank])
    # gradient averaging == gradient of the global-batch mean loss
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    target = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    lo, hi = mdist.shard_range(8, rank, world)
    red = mdist.GradAllReducer(model.parameters(), bucket_mb=0.0001)     # tiny buckets: several collectives
    torch.nn.functional.mse_loss(model(data[lo:hi]), target[lo:hi]).backward()
    red.reduce()
    got = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    torch.nn.functional.mse_loss(model(data), target).backward()
    want = [p.grad.clone() for p in model.parameters()]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, want))
    # overlapped mode: all-reduces launched from gradient hooks during backward, two steps (state resets), and a
    # parameter that takes no part in the loss (its bucket is completed by finish() with zeros)
    model.zero_grad(set_to_none=True)
    extra = torch.nn.Parameter(torch.ones(5))
    red2 = mdist.GradAllReducer(list(model.parameters()) + [extra], bucket_mb=0.0001, overlap=True)
    ok2, fired = True, 0
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        red2.begin_step()
        torch.nn.functional.mse_loss(model(data[lo:hi]), target[lo:hi]).backward()
        fired = red2.launched_in_backward
        red2.finish()
        got2 = [p.grad.clone() for p in model.parameters()]
        ok2 = ok2 and all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got2, want))
        ok2 = ok2 and extra.grad is not None and float(extra.grad.abs().max()) == 0.0
        extra.grad = None
    red2.close()
    q.put((rank, sizes, mx, ok and ok2, len(red.buckets), fired))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sizes, mx, ok, nb, fired in res:
        assert sizes == [(0, 6), (6, 11)]
        assert mx == [11.0, 5.0]
        assert ok, "bucketed gradient average differs from the global-batch gradient"
        assert nb > 1
        assert fired >= 1, "no bucket was all-reduced from a gradient hook during backward"


def test_shard_range_properties():
    from medmamba_b200.dist import shard_range
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
