"""BASELINE.json configs[4]: linear-probe training on frozen (cached) SigLIP features, data parallel, one process per GPU.

  python tools/bench_probe_train.py [--steps 200 --warmup 20]                                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29517 tools/bench_probe_train.py                                               # N GPUs

A step = one iteration of the reference's hot loop #2 (probing/train/base_trainer.py:166-181 with ``compute_losses`` of
singletask_trainer.py:183-191) on a batch of 256 cached feature rows PER GPU (weak scaling):
BatchNorm1d (batch statistics) -> Dropout(0.3) -> Linear(1152, 9) -> CrossEntropyLoss(weight, ignore_index=-1) ->
backward -> AdamW(lr 1e-4, wd 1e-4), the loss read back every step as the reference's ``_accumulate`` does.

Arms, all on the same cached features (synthetic: N(0.2, 1.3) rows, labels uniform with 5 % of -1):
  * ``value``        ProbeTrainStep (vlmfw_probe_* kernels, two NCCL all-reduces per step: BatchNorm sums [2D] fp64 and
                     the flat gradient bucket) replayed from one CUDA graph per step (``capture`` / ``step_graph``),
                     features already resident in HBM; CUDA events, max over ranks; ``eager`` = the same kernels
                     launched one by one (``step``);
  * ``e2e``          the same through host batches: pinned-host -> device copy of the rank's [256, 1152] fp32 rows and
                     labels inside the timed region, like the reference's ``x.to(device)`` of a DataLoader batch;
  * ``torch_gpu``    (rank 0, after the timed runs) the reference's loop as it stands — torch modules, fp16 autocast +
                     GradScaler (base_trainer.py:32,67-69), AdamW — on ONE GPU from host batches: the reference has no
                     multi-GPU training, so this is its best case;
  * ``cpu_baseline`` the same torch loop on the host cores.
Rank 0 prints one JSON line. Nothing here imports oracle/: the torch arms are baselines, written out below.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import _lib, parallel  # noqa: E402
from vlmfw_b200.probe_train import ProbeTrainStep  # noqa: E402

D, C, B_LOCAL, P_DROP = 1152, 9, 256, 0.3
# kernels of this library per step: bn_stats, dropout_mask, probe_train_fwd, probe_train_bwd, adamw, bn_running
# (csrc/probe.cu; every entry point launches exactly what its name says) — NCCL's two all-reduce kernels not counted
KERNELS_PER_STEP = 6


def torch_reference_loop(xs, ys, cw, device, steps, warmup, amp):
    """The reference's single-task loop on cached features, restated with its own building blocks (make_head of
    probing/models/base_probe.py:7-12; criterion singletask_trainer.py:105-108; optimiser :162-168)."""
    import torch.nn as nn

    torch.manual_seed(0)
    head = nn.Sequential(nn.BatchNorm1d(D), nn.Dropout(P_DROP), nn.Linear(D, C)).to(device).train()
    crit = nn.CrossEntropyLoss(weight=cw.to(device), ignore_index=-1)
    opt = torch.optim.AdamW(head.parameters(), lr=1e-4, weight_decay=1e-4)
    scaler = torch.amp.GradScaler(device="cuda", enabled=amp)
    total = 0.0

    def one(i):
        nonlocal total
        x = xs[i % len(xs)].to(device, non_blocking=True)
        y = ys[i % len(ys)].to(device, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        with torch.autocast(device_type=device.type, dtype=torch.float16, enabled=amp):
            loss = crit(head(x), y)
        if scaler.is_enabled():
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        else:
            loss.backward()
            opt.step()
        total += float(loss.item())  # base_trainer._accumulate reads the loss every step

    for i in range(warmup):
        one(i)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for i in range(steps):
        one(warmup + i)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    return steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--pool", type=int, default=32, help="distinct cached batches cycled through")
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time only the eager launches")
    args = ap.parse_args()

    env = parallel.init_distributed("nccl")
    dev = torch.device("cuda", env.local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    nb = B_LOCAL * env.world
    lo, hi = parallel.shard_range(nb, env.rank, env.world)
    g = torch.Generator().manual_seed(11)
    # every rank draws the same global pool and keeps its slice: the cached `{split}_features.pt` rows of this rank
    xs_host, ys_host = [], []
    for _ in range(args.pool):
        x = torch.randn(nb, D, generator=g) * 1.3 + 0.2
        y = torch.randint(0, C, (nb,), generator=g)
        y[torch.rand(nb, generator=g) < 0.05] = -1
        xs_host.append(x[lo:hi].contiguous().pin_memory())
        ys_host.append(y[lo:hi].contiguous().pin_memory())
    cw = torch.rand(C, generator=g) + 0.5
    xs_dev = [x.to(dev) for x in xs_host]
    ys_dev = [y.to(dev) for y in ys_host]

    def timed(resident: bool, use_graph: bool):
        ts = ProbeTrainStep(D, C, device=dev, class_weight=cw, dropout_p=P_DROP, seed=3)
        if use_graph:
            ts.capture(hi - lo, n_global=nb, global_offset=lo)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loss = 0.0

        def one(i):
            j = i % args.pool
            x, y = (xs_dev[j], ys_dev[j]) if resident else (xs_host[j], ys_host[j])
            if use_graph:
                return ts.step_graph(x, y)  # stages x / y (device or pinned host) into the graph's input buffers
            if not resident:
                x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            return ts.step(x, y, n_global=nb, global_offset=lo)

        for i in range(args.warmup):
            one(i)
        torch.cuda.synchronize(dev)
        if env.is_distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(args.steps):
            loss = one(args.warmup + i)
        e1.record()
        torch.cuda.synchronize(dev)
        if env.is_distributed:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if env.is_distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), loss, ts

    ms_eager, loss_eager, ts = timed(True, False)
    ms_eager_e2e, _, _ = timed(False, False)
    graph_err = None
    graphed = []
    try:
        if args.no_graph:
            raise RuntimeError("--no-graph")
        ms_res, loss_res, g1 = timed(True, True)
        graphed.append(g1)
        ms_e2e, loss_e2e, g2 = timed(False, True)
        graphed.append(g2)
    except Exception as e:  # capture of the NCCL collectives refused: report the eager launches of the same kernels
        if env.is_distributed and not args.no_graph:
            raise  # ranks must not diverge
        graph_err = repr(e)
        print(f"[bench_probe_train] CUDA graph capture failed ({e}); reporting eager launches", file=sys.stderr)
        ms_res, loss_res, ms_e2e, loss_e2e = ms_eager, loss_eager, ms_eager_e2e, loss_eager
    launches = (KERNELS_PER_STEP + 1) * args.steps  # + the step-counter kernel of the graph form

    if env.rank == 0:
        res = {
            "metric": "probe train steps/sec on cached SigLIP features", "unit": "steps/s",
            "value": round(args.steps / (ms_res / 1e3), 1), "n_gpus": env.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "samples_per_s": round(args.steps * nb / (ms_res / 1e3), 0),
            "config": {"workload": "LinearProbe head (BatchNorm1d -> Dropout 0.3 -> Linear(1152, 9)), class-weighted CE "
                                   "with ignore_index -1, AdamW(1e-4, wd 1e-4), batch 256 rows per GPU",
                       "global_batch": nb, "parallelism": f"dp{env.world}",
                       "exchange": "2 NCCL all-reduces per step (BatchNorm sums, gradient bucket)",
                       "allreduce_bytes_per_step": 2 * D * 8 + (ts.n_params + 2) * 4,
                       "loss_read_back_every_step": True, "cuda_graph": graph_err is None},
            "eager": {"value": round(args.steps / (ms_eager / 1e3), 1), "e2e": round(args.steps / (ms_eager_e2e / 1e3), 1),
                      "unit": "steps/s", "what": "the same kernels launched one by one (ProbeTrainStep.step)",
                      "final_loss": round(loss_eager, 6)},
            "e2e": {"value": round(args.steps / (ms_e2e / 1e3), 1), "unit": "steps/s",
                    "h2d_bytes_per_step": B_LOCAL * D * 4 + B_LOCAL * 8, "d2h_bytes_per_step": 8,
                    "api": "ProbeTrainStep.step_graph(pinned host batch)"},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "final_loss": round(loss_res, 6), "final_loss_e2e": round(loss_e2e, 6),
        }
        if not args.no_baselines:
            # the reference's loop sees the GLOBAL batch on its single device: same samples per step as the DP job
            gx = torch.Generator().manual_seed(11)
            xs_full, ys_full = [], []
            for _ in range(min(args.pool, 8)):
                x = torch.randn(nb, D, generator=gx) * 1.3 + 0.2
                y = torch.randint(0, C, (nb,), generator=gx)
                y[torch.rand(nb, generator=gx) < 0.05] = -1
                xs_full.append(x.pin_memory())
                ys_full.append(y.pin_memory())
            sps = torch_reference_loop(xs_full, ys_full, cw, dev, args.steps, args.warmup, amp=True)
            res["torch_gpu"] = {"value": round(sps, 1), "unit": "steps/s", "gpus": 1, "batch": nb,
                                "what": "reference loop (torch modules, fp16 autocast + GradScaler, AdamW) on one GPU, "
                                        "global batch per step, host batches"}
        if not args.no_baselines and not env.is_distributed:  # host-core baseline at N = 1 only (torchrun pins OMP to 1 thread)
            torch.set_num_threads(os.cpu_count() or 1)
            xs_c = [x.clone() for x in xs_full]
            ys_c = [y.clone() for y in ys_full]
            sps_cpu = torch_reference_loop(xs_c, ys_c, cw, torch.device("cpu"), min(args.steps, 50), 5, amp=False)
            res["cpu_baseline"] = {"value": round(sps_cpu, 1), "unit": "steps/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "sample": f"same loop on the host cores, global batch {nb} per step, "
                                                              f"{min(args.steps, 50)} timed steps"}
        print(json.dumps(res), flush=True)
    if env.is_distributed:
        # The N = 2 run of this script printed its line and then sat in the teardown until `timeout` killed it: the
        # process group was being destroyed while CUDA graphs holding captured NCCL kernels were still alive. Release
        # the graphs first, and leave without the communicator teardown (the line is out, nothing is left to flush).
        for t in graphed:
            t.release_graph()
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
