"""Native probe train step on cached features (linear head), data parallel.

Restates one iteration of the reference's hot loop #2 — probing/train/base_trainer.py:166-181 with
``compute_losses`` of probing/train/singletask_trainer.py:183-191: ``head_only(x)`` = BatchNorm1d (train: batch
statistics) -> Dropout(p) -> Linear, ``CrossEntropyLoss(weight=w_c, ignore_index=-1)`` (:105-108), backward, AdamW
(lr 1e-4, wd 1e-4; :162-168) — as three CUDA phases whose outputs are plain sums, so that sharding the batch over
ranks and all-reducing between the phases lands on exactly the single-process step for the concatenated batch:

    1. vlmfw_probe_bn_stats      local sum x, sum x^2 (fp64)            -> all_reduce(SUM)   [2D] doubles
    2. vlmfw_probe_linear_grads  fwd + bwd with GLOBAL batch statistics -> all_reduce(SUM)   [C*D + C + 2D + 2] floats
    3. vlmfw_probe_adamw + vlmfw_probe_bn_running_update (identical on every rank)

The reference's fp16 autocast + GradScaler (base_trainer.py:32,67-69) is a no-op for the result at these magnitudes
and is not reproduced: everything here is fp32. Dropout masks are keyed by (seed, step, GLOBAL row) so the sharded and
the single-process step draw the same mask.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .parallel import allreduce_sum_, shard_range


class ProbeTrainStep:
    def __init__(self, dim: int, classes: int, device="cuda", lr: float = 1e-4, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, dropout_p: float = 0.3,
                 class_weight: Optional[torch.Tensor] = None, bn_eps: float = 1e-5, bn_momentum: float = 0.1,
                 seed: int = 0, group=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ProbeTrainStep runs on CUDA only: vlmfw_b200 has no CPU fallback")
        self.dim, self.classes = dim, classes
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.dropout_p, self.bn_eps, self.bn_momentum = dropout_p, bn_eps, bn_momentum
        self.seed, self.group = seed, group
        self.n_params = classes * dim + classes + 2 * dim
        dev = self.device
        # flat parameter vector laid out like the gradient bucket: [W (C*D) | b (C) | gamma (D) | beta (D)]
        self.params = torch.zeros(self.n_params, device=dev)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.running_mean = torch.zeros(dim, device=dev)
        self.running_var = torch.ones(dim, device=dev)
        self.num_batches_tracked = 0
        self.step_count = 0
        self.class_weight = None if class_weight is None else class_weight.to(dev, torch.float32).contiguous()
        self.stats = torch.zeros(2 * dim, dtype=torch.float64, device=dev)
        self.bucket = torch.zeros(self.n_params + 2, device=dev)
        self._lib = _lib.load()
        # nn.Linear / BatchNorm1d default init, seeded
        g = torch.Generator().manual_seed(seed)
        bound = 1.0 / dim ** 0.5
        self.weight.copy_((torch.rand(classes, dim, generator=g) * 2 - 1) * bound)
        self.bias.copy_((torch.rand(classes, generator=g) * 2 - 1) * bound)
        self.gamma.fill_(1.0)

    # views into the flat vector
    @property
    def weight(self):
        return self.params[: self.classes * self.dim].view(self.classes, self.dim)

    @property
    def bias(self):
        o = self.classes * self.dim
        return self.params[o: o + self.classes]

    @property
    def gamma(self):
        o = self.classes * self.dim + self.classes
        return self.params[o: o + self.dim]

    @property
    def beta(self):
        o = self.classes * self.dim + self.classes + self.dim
        return self.params[o: o + self.dim]

    def load_head_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        """Reference keys (nn.Sequential indices of make_head): ``0.*`` BatchNorm1d, ``2.*`` Linear."""
        self.gamma.copy_(sd[prefix + "0.weight"])
        self.beta.copy_(sd[prefix + "0.bias"])
        self.running_mean.copy_(sd[prefix + "0.running_mean"])
        self.running_var.copy_(sd[prefix + "0.running_var"])
        self.num_batches_tracked = int(sd.get(prefix + "0.num_batches_tracked", 0))
        self.weight.copy_(sd[prefix + "2.weight"])
        self.bias.copy_(sd[prefix + "2.bias"])

    def head_state_dict(self, prefix: str = "") -> Dict[str, torch.Tensor]:
        return {
            prefix + "0.weight": self.gamma.clone(), prefix + "0.bias": self.beta.clone(),
            prefix + "0.running_mean": self.running_mean.clone(), prefix + "0.running_var": self.running_var.clone(),
            prefix + "0.num_batches_tracked": torch.tensor(self.num_batches_tracked),
            prefix + "2.weight": self.weight.clone(), prefix + "2.bias": self.bias.clone(),
        }

    def dropout_keep_mask(self, n_global: int, lo: int, hi: int) -> Optional[torch.Tensor]:
        """uint8 [hi-lo, D]; row r of the GLOBAL batch always gets the same mask regardless of the sharding."""
        if self.dropout_p <= 0.0:
            return None
        g = torch.Generator(device="cpu").manual_seed(self.seed * 1000003 + self.step_count)
        full = torch.rand(n_global, self.dim, generator=g) >= self.dropout_p
        return full[lo:hi].to(torch.uint8).to(self.device).contiguous()

    # ---- the three phases (exposed separately so tests can drive emulated ranks on one GPU) ----------------------
    def _prep(self, x, y):
        if not x.is_cuda:
            raise RuntimeError("features must be CUDA tensors")
        return x.to(torch.float32).contiguous(), y.to(device=x.device, dtype=torch.int64).contiguous()

    def phase1_stats(self, x: torch.Tensor) -> torch.Tensor:
        """Local BatchNorm sums into ``self.stats`` (fp64 [2D]); the caller all-reduces it."""
        with torch.cuda.device(x.device):
            _lib.check(self._lib.vlmfw_probe_bn_stats(x.data_ptr(), x.shape[0], self.dim, self.stats.data_ptr(),
                                                      torch.cuda.current_stream(x.device).cuda_stream))
        return self.stats

    def phase2_grads(self, x: torch.Tensor, y: torch.Tensor, n_global: int,
                     keep_mask: Optional[torch.Tensor]) -> torch.Tensor:
        """Forward + backward of the local rows against the GLOBAL statistics in ``self.stats`` into ``self.bucket``."""
        b = x.shape[0]
        self.last_logits = torch.empty(b, self.classes, device=x.device)
        scratch = torch.empty(b * (self.classes + 2), device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.vlmfw_probe_linear_grads(
                x.data_ptr(), y.data_ptr(), b, self.dim, self.classes, self.stats.data_ptr(), n_global, self.bn_eps,
                self.gamma.data_ptr(), self.beta.data_ptr(), self.weight.data_ptr(), self.bias.data_ptr(),
                None if self.class_weight is None else self.class_weight.data_ptr(),
                None if keep_mask is None else keep_mask.data_ptr(), self.dropout_p, self.bucket.data_ptr(),
                self.last_logits.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
        return self.bucket

    def phase3_update(self, n_global: int) -> float:
        """AdamW on the flat parameters from the (all-reduced) bucket + BatchNorm running statistics."""
        self.step_count += 1
        dev = self.params.device
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(self._lib.vlmfw_probe_adamw(
                self.params.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.bucket.data_ptr(),
                self.n_params, self.n_params + 1, self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                self.wd, st))
            _lib.check(self._lib.vlmfw_probe_bn_running_update(
                self.running_mean.data_ptr(), self.running_var.data_ptr(), self.stats.data_ptr(), self.dim, n_global,
                self.bn_momentum, st))
        self.num_batches_tracked += 1
        tail = self.bucket[self.n_params:].tolist()  # [loss_sum, weight_sum] — the step's only D2H read
        return tail[0] / tail[1] if tail[1] > 0 else 0.0

    def step(self, x: torch.Tensor, y: torch.Tensor, n_global: Optional[int] = None, global_offset: int = 0,
             keep_mask: Optional[torch.Tensor] = None) -> float:
        """One optimisation step on this rank's slice ``x [B_local, D]`` fp32, ``y [B_local]`` int64 (-1 = ignore).
        Returns the global mean loss (as CrossEntropyLoss would report on the concatenated batch)."""
        x, y = self._prep(x, y)
        b = x.shape[0]
        n_global = b if n_global is None else n_global
        if keep_mask is None:
            keep_mask = self.dropout_keep_mask(n_global, global_offset, global_offset + b)
        allreduce_sum_(self.phase1_stats(x), self.group)
        allreduce_sum_(self.phase2_grads(x, y, n_global, keep_mask), self.group)
        return self.phase3_update(n_global)
