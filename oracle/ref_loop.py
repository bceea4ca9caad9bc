"""ORACLE (test infrastructure / reported CPU baseline): the reference's own loops restated with torch modules.

* ``build_reference_heads`` — eval-mode probe heads exactly as the reference builds them (``make_head`` of
  probing/models/base_probe.py:7-12: BatchNorm1d -> Dropout -> Linear), loaded with the seeded parameters the benchmark
  uses; ``bench.py --impl reference`` and its ``cpu_baseline`` leg put them on top of the live HF tower (hf_ref.py).
* ``reference_train_loop`` — the single-task loop on cached features as the reference runs it: ``head_only(x)``,
  ``CrossEntropyLoss(weight, ignore_index=-1)`` (probing/train/singletask_trainer.py:105-108,183-191), fp16 autocast +
  GradScaler on a GPU (probing/train/base_trainer.py:32,67-69), AdamW(lr 1e-4, wd 1e-4) (singletask_trainer.py:162-168),
  the loss read back every step (base_trainer._accumulate). Timed as the CPU baseline of ``bench.py --workload
  probe_train`` and as its ``--impl reference`` arm; never on the product path.
"""
from __future__ import annotations

import time
from typing import Dict, Mapping

import torch
import torch.nn as nn


def make_head(in_dim: int, out_dim: int, dropout_p: float) -> nn.Sequential:
    """probing/models/base_probe.py:7-12."""
    return nn.Sequential(nn.BatchNorm1d(in_dim), nn.Dropout(p=dropout_p), nn.Linear(in_dim, out_dim))


def build_reference_heads(dim: int, tasks: Mapping[str, int], params: Mapping[str, Mapping[str, torch.Tensor]],
                          dropout_p: float = 0.3) -> Dict[str, nn.Sequential]:
    heads = {}
    for t, c in tasks.items():
        h = make_head(dim, c, dropout_p).eval()
        h.load_state_dict(dict(params[t]))
        heads[t] = h
    return heads


def reference_train_loop(xs, ys, class_weight, device, steps: int, warmup: int, amp: bool, dim: int, classes: int,
                         dropout_p: float = 0.3) -> float:
    """steps/s of the reference's loop over the given (host) batches on ``device``."""
    torch.manual_seed(0)
    head = make_head(dim, classes, dropout_p).to(device).train()
    crit = nn.CrossEntropyLoss(weight=class_weight.to(device), ignore_index=-1)
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
        total += float(loss.item())

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
