"""Multitask / deeper-head probe training on cached CLIP features (BASELINE configs[2]'s 5-head MiviaPar probe;
reference probing/train/multitask_trainer.py:188-263 over base_trainer.py:166-181), one process per GPU.

  python tools/bench_probe_multitask.py [--deeper] [--steps 500 --warmup 30]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29518 \
      tools/bench_probe_multitask.py [--deeper]

A step = shared BatchNorm sums -> (all-reduce) -> every head's forward / backward on 256 cached feature rows per GPU ->
(all-reduce of ONE flat bucket) -> AdamW per head, per-task losses read back, as the reference's loop does. Arms:
``value`` = MultiTaskProbeTrainStep.step_graph (one CUDA graph per step), ``eager`` = the same kernels launched one by one,
``torch_gpu`` (rank 0) = the reference's loop restated with torch modules on one GPU. Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import _lib, parallel  # noqa: E402
from vlmfw_b200.probe_train import MultiTaskProbeTrainStep  # noqa: E402

D, B_LOCAL, P_DROP, HIDDEN = 1024, 256, 0.3, 512
TASKS = {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}


def torch_loop(xs, ys, device, steps, warmup, deeper):
    """make_head / make_head_deeper per task (probing/models/base_probe.py:7-22), CE with ignore_index, sum of task losses,
    one AdamW — fp32, no autocast (the heads are tiny)."""
    import torch.nn as nn

    torch.manual_seed(0)

    def head(c):
        if deeper:
            return nn.Sequential(nn.BatchNorm1d(D), nn.Dropout(P_DROP), nn.Linear(D, HIDDEN), nn.GELU(), nn.Dropout(P_DROP),
                                 nn.Linear(HIDDEN, c))
        return nn.Sequential(nn.BatchNorm1d(D), nn.Dropout(P_DROP), nn.Linear(D, c))

    heads = nn.ModuleDict({t: head(c) for t, c in TASKS.items()}).to(device).train()
    crit = nn.CrossEntropyLoss(ignore_index=-1)
    opt = torch.optim.AdamW(heads.parameters(), lr=1e-4, weight_decay=1e-4)

    def one(i):
        x, y = xs[i % len(xs)], ys[i % len(ys)]
        opt.zero_grad(set_to_none=True)
        losses = [crit(heads[t](x), y[:, k]) for k, t in enumerate(TASKS)]
        total = sum(losses)
        total.backward()
        opt.step()
        return [float(v) for v in torch.stack(losses).tolist()]

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(warmup + i)
    e1.record()
    torch.cuda.synchronize(device)
    return steps / (e0.elapsed_time(e1) / 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--deeper", action="store_true")
    args = ap.parse_args()
    env = parallel.init_distributed("nccl")
    dev = torch.device("cuda", env.local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    nb = B_LOCAL * env.world
    lo, hi = parallel.shard_range(nb, env.rank, env.world)
    g = torch.Generator().manual_seed(9)
    pool = 16
    xs = [(torch.randn(nb, D, generator=g) * 1.3 + 0.2)[lo:hi].to(dev) for _ in range(pool)]
    ys = [torch.stack([torch.randint(-1, c, (nb,), generator=g) for c in TASKS.values()], dim=1)[lo:hi].to(dev)
          for _ in range(pool)]

    def make():
        return MultiTaskProbeTrainStep(D, TASKS, dev, deeper=args.deeper, hidden_dim=HIDDEN, dropout_p=P_DROP, seed=5)

    def timed(fn):
        for i in range(args.warmup):
            fn(i)
        torch.cuda.synchronize(dev)
        if env.is_distributed:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fn(args.warmup + i)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if env.is_distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return args.steps / (float(ms.item()) / 1e3)

    ts_g = make()
    ts_g.capture(hi - lo, n_global=nb, global_offset=lo)
    graph_rate = timed(lambda i: ts_g.step_graph(xs[i % pool], ys[i % pool]))
    ts_e = make()
    eager_rate = timed(lambda i: ts_e.step(xs[i % pool], ys[i % pool], n_global=nb, global_offset=lo))
    ts_g.release_graph()
    line = None
    if env.rank == 0:
        torch_rate = torch_loop(xs, ys, dev, min(args.steps, 200), 10, args.deeper) if env.world == 1 else None
        line = {"metric": "multitask probe train steps/sec on cached CLIP features", "value": round(graph_rate, 1),
                "unit": "steps/s", "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 / graph_rate, 4), "eager_steps_per_s": round(eager_rate, 1),
                "torch_gpu_steps_per_s": None if torch_rate is None else round(torch_rate, 1), "dtype": "f32",
                "data": "synthetic", "gpu_launches_per_step": ts_g.kernels_per_step(),
                "config": {"workload": f"MultiTaskProbe ({'make_head_deeper, hidden 512' if args.deeper else 'make_head'}) x 5 MiviaPar "
                                       f"tasks {TASKS} on cached CLIP ViT-L features (D = 1024), batch 256 rows per GPU, per-task "
                                       "losses read back every step", "global_batch": nb, "parallelism": f"dp{env.world}",
                           "cuda_graph": True}}
    if env.is_distributed:
        torch.cuda.synchronize(dev)
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
