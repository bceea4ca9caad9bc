"""Multi-GPU check of the hot path, one process per GPU over NCCL (BASELINE.json configs[1] sharding and configs[4]):

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

1. inference: the global batch is sharded in contiguous slices (parallel.shard_range), every rank runs tower + probe on
   its slice, labels / features are gathered in order (parallel.all_gather_rows) and compared on rank 0 with the same
   global batch run on one GPU — must be bit-identical (no data-path collective, same kernels, same per-image work);
2. probe training (linear head on cached features): K steps with the batch sharded over N ranks and the two NCCL
   all-reduces per step (BatchNorm sums [2D] fp64, gradient bucket fp32) vs the single-process step on the whole batch.
Rank 0 prints one JSON line; exit code 1 on a mismatch."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import parallel, probes, specs, weights  # noqa: E402
from vlmfw_b200.probe_train import ProbeTrainStep  # noqa: E402
from vlmfw_b200.vision_backbone import backbone_from_spec  # noqa: E402


def main():
    env = parallel.init_distributed("nccl")
    dev = torch.device("cuda", env.local_rank)
    torch.cuda.set_device(dev)
    spec = specs.SPECS[os.environ.get("DP_CHECK_TOWER", "siglip_so400m")]
    per_rank = int(os.environ.get("DP_CHECK_BATCH", "32"))
    n = per_rank * env.world + (3 if env.world > 1 else 0)  # ragged on purpose: the first ranks get one row more
    bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, 0), device=dev)
    probe = probes.LinearProbe(bb, 9).to(dev).eval()
    probe.classifier.load_state_dict(weights.seeded_head_params(spec.hidden, 9, 0, seed=100))
    imgs = weights.seeded_images_u8(n, spec.image_size, 7)  # same global batch on every rank
    lo, hi = parallel.shard_range(n, env.rank, env.world)
    mine = imgs[lo:hi].to(dev)
    feats = probe.extract_features(mine)
    labels = probe.predict(mine)
    all_feats = parallel.all_gather_rows(feats, n)
    all_labels = parallel.all_gather_rows(labels.view(-1, 1), n).view(-1)
    res = {"world": env.world, "tower": spec.name, "global_batch": n}
    ok = True
    if env.rank == 0:
        full = imgs.to(dev)
        ref_feats = torch.cat([probe.extract_features(full[i:i + per_rank]) for i in range(0, n, per_rank)])
        ref_labels = torch.cat([probe.predict(full[i:i + per_rank]) for i in range(0, n, per_rank)])
        res["features_bit_identical"] = bool(torch.equal(all_feats, ref_feats))
        res["labels_identical"] = bool(torch.equal(all_labels, ref_labels))
        ok &= res["features_bit_identical"] and res["labels_identical"]

    # ---- probe training on cached features
    d, c, b_local, steps = spec.hidden, 9, 256, 20
    g = torch.Generator().manual_seed(11)
    nb = b_local * env.world
    xs = torch.randn(steps, nb, d, generator=g) * 1.3 + 0.2
    ys = torch.randint(0, c, (steps, nb), generator=g)
    ys[:, ::17] = -1
    cw = torch.rand(c, generator=g) + 0.5
    dp = ProbeTrainStep(d, c, device=dev, class_weight=cw, seed=3)
    tlo, thi = parallel.shard_range(nb, env.rank, env.world)
    torch.cuda.synchronize(dev)
    dist.barrier() if env.is_distributed else None
    t0 = time.perf_counter()
    dp_losses = [dp.step(xs[s, tlo:thi].to(dev), ys[s, tlo:thi].to(dev), n_global=nb, global_offset=tlo) for s in range(steps)]
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    # the graph form of the same steps, once per exchange: fused NVLink peer-memory slots (p2p) and NCCL all-reduces
    graph_runs = {}
    for comm in ("p2p", "nccl"):
        gs = ProbeTrainStep(d, c, device=dev, class_weight=cw, seed=3, comm=comm)
        gs.capture(thi - tlo, n_global=nb, global_offset=tlo)
        losses = [gs.step_graph(xs[s, tlo:thi].to(dev), ys[s, tlo:thi].to(dev)) for s in range(steps)]
        bad = gs.comm_error()
        # every rank must hold bit-identical parameters (the slots are added in rank order on every rank)
        mine = gs.params.clone()
        ref0 = mine.clone()
        if env.is_distributed:
            dist.broadcast(ref0, 0)
        same = torch.tensor([1 if torch.equal(mine, ref0) else 0], device=dev)
        if env.is_distributed:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        graph_runs[comm] = (gs.params.clone(), losses, bool(same.item()), bad)
        gs.release_graph()
    if env.rank == 0:
        single = ProbeTrainStep(d, c, device=dev, class_weight=cw, seed=3, group=False)
        single.group = None
        # single-process reference: same global batch, no sharding (collectives are no-ops on a 1-rank group)
        saved = dist.is_initialized()
        one_losses = []
        for s in range(steps):
            x, y = single._prep(xs[s].to(dev), ys[s].to(dev))
            keep = single.dropout_keep_mask(nb, 0, nb)
            single.phase1_stats(x)
            single.phase2_grads(x, y, nb, keep)
            one_losses.append(single.phase3_update(nb))
        perr = (dp.params - single.params).abs().max().item()
        lerr = max(abs(a - b) for a, b in zip(dp_losses, one_losses))
        res.update({"train_steps": steps, "train_batch_per_gpu": b_local, "param_max_abs_diff_vs_single_process": perr,
                    "loss_max_abs_diff": lerr, "train_steps_per_s": round(steps / dt, 1),
                    "allreduce_bytes_per_step": 2 * d * 8 + (dp.n_params + 2) * 4})
        ok &= perr <= 1e-5 and lerr <= 1e-5
        for comm, (gp, gl, same, bad) in graph_runs.items():
            gerr = (gp - single.params).abs().max().item()
            glerr = max(abs(a - b) for a, b in zip(gl, one_losses))
            res[f"graph_{comm}"] = {"param_max_abs_diff_vs_single_process": gerr, "loss_max_abs_diff": glerr,
                                    "params_bit_identical_across_ranks": same, "comm_error": bad}
            ok &= gerr <= 1e-5 and glerr <= 1e-5 and same and not bad
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    if env.is_distributed:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
