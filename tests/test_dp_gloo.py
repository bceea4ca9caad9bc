"""CPU, world_size 2, gloo: the data-parallel plumbing of the path (sharding, ordered gather, the two all-reduces of the
probe train step) — the collectives are the same calls NCCL serves on the GPUs."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from oracle import probe_oracle, tower_oracle as O
    from vlmfw_b200 import parallel

    env = parallel.init_distributed("gloo")
    assert env.rank == rank and env.world == world
    torch.manual_seed(0)  # same data on every rank, each takes its slice
    n, d, c, p = 37, 32, 5, 0.3
    x = torch.randn(n, d) * 1.5 + 0.3
    y = torch.randint(0, c, (n,))
    y[::5] = -1
    keep = (torch.rand(n, d) >= p).to(torch.uint8)
    cw = torch.rand(c) + 0.5
    state = O.seeded_head_params(d, c, 0, seed=2)
    lo, hi = parallel.shard_range(n, rank, world)
    # ordered gather of per-rank feature blocks
    feats = x[lo:hi] * 2
    full = parallel.all_gather_rows(feats, n)
    ok_gather = torch.equal(full, x * 2)
    # the two exchange steps of the train step
    stats = parallel.allreduce_sum_(probe_oracle.phase1_stats(x[lo:hi]))
    bucket = parallel.allreduce_sum_(probe_oracle.phase2_bucket(
        x[lo:hi], y[lo:hi], stats, n, state["0.weight"], state["0.bias"], state["2.weight"], state["2.bias"], cw,
        keep[lo:hi], p))
    single_stats = probe_oracle.phase1_stats(x)
    single = probe_oracle.phase2_bucket(x, y, single_stats, n, state["0.weight"], state["0.bias"], state["2.weight"],
                                        state["2.bias"], cw, keep, p)
    ok_grad = torch.allclose(bucket, single, rtol=1e-12, atol=1e-12)
    q.put((rank, ok_gather, ok_grad, float((bucket - single).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, ok_grad, err in res:
        assert ok_gather, f"rank {rank}: gathered rows out of order"
        assert ok_grad, f"rank {rank}: all-reduced gradient differs from single-process gradient by {err}"
