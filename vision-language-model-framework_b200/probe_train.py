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
and is not reproduced: everything here is fp32. Dropout masks come from a counter-based generator on the device keyed by
(seed, step, task, GLOBAL row, column) — vlmfw_dropout_keep_mask — so the sharded and the single-process step draw the
same mask and every rank draws only its own rows.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .parallel import allreduce_sum_, shard_range


def mask_key(seed: int, step: int, task_index: int = 0) -> int:
    """64-bit key of one dropout draw: (seed, optimisation step, task); the second dropout of a deeper head uses key + 1."""
    return (((seed * 1000003 + step) * 131 + task_index) * 2) & 0xFFFFFFFFFFFFFFFF


MASK_KEY_STEP_STRIDE = 262  # mask_key(seed, step + 1, t) - mask_key(seed, step, t): what one step adds to the key


def device_keep_mask(lib, device, rows: int, cols: int, row0: int, key: int, p: float) -> torch.Tensor:
    """uint8 [rows, cols] keep mask of global rows [row0, row0 + rows), drawn on the device (vlmfw_dropout_keep_mask):
    nothing crosses PCIe and no rank draws more than its own rows."""
    keep = torch.empty(rows, cols, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.vlmfw_dropout_keep_mask(keep.data_ptr(), rows, cols, row0, key & 0xFFFFFFFFFFFFFFFF, float(p),
                                               torch.cuda.current_stream(device).cuda_stream))
    return keep


class ProbeTrainStep:
    def __init__(self, dim: int, classes: int, device="cuda", lr: float = 1e-4, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, dropout_p: float = 0.3,
                 class_weight: Optional[torch.Tensor] = None, bn_eps: float = 1e-5, bn_momentum: float = 0.1,
                 seed: int = 0, group=None, comm: str = "auto"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ProbeTrainStep runs on CUDA only: vlmfw_b200 has no CPU fallback")
        if comm not in ("auto", "nccl", "p2p"):
            raise ValueError("comm must be 'auto', 'nccl' or 'p2p'")
        # exchange of the graph-form step: "p2p" = fused into the step's kernels over NVLink peer memory (vlmfw_comm, one node),
        # "nccl" = two ncclAllReduce calls between the kernels; "auto" = p2p when a process group is up, else nothing to exchange
        self.comm = comm
        self._comm_h = None
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
        self._graph = None  # set by capture()
        self._lib = _lib.load()
        # nn.Linear / BatchNorm1d default init, seeded
        g = torch.Generator().manual_seed(seed)
        bound = 1.0 / dim ** 0.5
        self.weight.copy_((torch.rand(classes, dim, generator=g) * 2 - 1) * bound)
        self.bias.copy_((torch.rand(classes, generator=g) * 2 - 1) * bound)
        self.gamma.fill_(1.0)

    # kernels of this library per graph-form step: bn_stats, dropout_mask, probe_train_fwd, probe_train_bwd, adamw,
    # bn_running, counter_add (csrc/probe.cu); NCCL's all-reduce kernels are not counted
    kernels_per_step = 7

    def _use_p2p(self) -> bool:
        import torch.distributed as dist

        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        return self.comm == "p2p" or (self.comm == "auto" and distributed)

    def exchange_description(self) -> str:
        if self._comm_h is not None:
            return ("fused into the step's kernels over NVLink peer memory: bn_stats / backward store every sum together with the "
                    "exchange's flag (one 8- / 16-byte store) into every rank's slots, the mask / AdamW kernels poll their local "
                    "slots and add them in rank order (BatchNorm sums [2D] fp64, gradient bucket [C*D + C + 2D + 2] fp32); no "
                    "fence, no collective kernel")
        return ("2 dependent NCCL all-reduces per step: BatchNorm batch sums [2D] fp64, then the flat gradient bucket "
                "[C*D + C + 2D + 2] fp32")

    def exchange_report(self, steps: int = 200) -> dict:
        """Latency of the NCCL form of the exchange measured in isolation (the baseline the fused form replaces): ``steps``
        DEPENDENT all-reduces of each of the two buffers, captured in one CUDA graph like the train step, timed with CUDA
        events (max over ranks)."""
        import torch.distributed as dist

        dev = self.device
        out = {}
        for name, buf in (("bn_sums_f64", torch.zeros_like(self.stats)), ("grad_bucket_f32", torch.zeros_like(self.bucket))):
            allreduce_sum_(buf, self.group)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(steps):
                    allreduce_sum_(buf, self.group)
            g.replay()
            torch.cuda.synchronize(dev)
            dist.barrier(self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=self.group)
            out[name] = {"bytes": buf.numel() * buf.element_size(), "us_per_allreduce": round(float(ms.item()) * 1e3 / steps, 2)}
            del g
        out["method"] = f"{steps} dependent ncclAllReduce calls replayed from one CUDA graph, CUDA events, max over ranks"
        return out

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
        return device_keep_mask(self._lib, self.device, hi - lo, self.dim, lo, mask_key(self.seed, self.step_count),
                                self.dropout_p)

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

    # ---- the whole step as ONE CUDA graph -----------------------------------------------------------------------
    def capture(self, b_local: int, n_global: Optional[int] = None, global_offset: int = 0) -> None:
        """Capture one optimisation step for a fixed local batch size into a CUDA graph: the six kernels, the two
        all-reduces (NCCL collectives are capturable) and the read-back of [loss_sum, weight_sum] into pinned host
        memory. Nothing in the launch sequence depends on the step number: AdamW's bias corrections and the per-step
        part of the dropout key are read from a device-side counter (vlmfw_probe_adamw_dev,
        vlmfw_dropout_keep_mask_dev, vlmfw_counter_add). Replaying costs one launch instead of seven plus their
        host-side argument marshalling: 13.5 k against 8.5 k steps/s at 256 rows on one B200.
        ``step_graph`` then runs the same arithmetic as ``step`` (same kernels, same masks); the eager ``step`` stays
        usable in between."""
        import torch.distributed as dist

        dev, d, c = self.device, self.dim, self.classes
        n_global = b_local if n_global is None else n_global
        self._g_shape = (b_local, n_global, global_offset)
        self._g_x = torch.zeros(b_local, d, device=dev)
        self._g_y = torch.full((b_local,), -1, dtype=torch.int64, device=dev)
        self._g_keep = torch.empty(b_local, d, dtype=torch.uint8, device=dev) if self.dropout_p > 0.0 else None
        self._g_logits = torch.empty(b_local, c, device=dev)
        self._g_scratch = torch.empty(b_local * (c + 2), device=dev)
        self._g_steps_done = torch.tensor([self.step_count], dtype=torch.int64, device=dev)
        self._g_steps_done_value = self.step_count
        self._g_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self._use_p2p():
            self._capture_p2p(b_local, n_global, global_offset)
            return
        if distributed:  # communicator set-up cannot be captured: run one collective of each dtype eagerly first
            allreduce_sum_(torch.zeros(2, dtype=torch.float64, device=dev), self.group)
            allreduce_sum_(torch.zeros(2, device=dev), self.group)
        torch.cuda.synchronize(dev)
        lib, key0 = self._lib, mask_key(self.seed, 0)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.cuda.graph(graph):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.vlmfw_probe_bn_stats(self._g_x.data_ptr(), b_local, d, self.stats.data_ptr(), st))
            allreduce_sum_(self.stats, self.group)
            if self._g_keep is not None:
                _lib.check(lib.vlmfw_dropout_keep_mask_dev(self._g_keep.data_ptr(), b_local, d, global_offset, key0,
                                                           self._g_steps_done.data_ptr(), MASK_KEY_STEP_STRIDE,
                                                           float(self.dropout_p), st))
            _lib.check(lib.vlmfw_probe_linear_grads(
                self._g_x.data_ptr(), self._g_y.data_ptr(), b_local, d, c, self.stats.data_ptr(), n_global, self.bn_eps,
                self.gamma.data_ptr(), self.beta.data_ptr(), self.weight.data_ptr(), self.bias.data_ptr(),
                None if self.class_weight is None else self.class_weight.data_ptr(),
                None if self._g_keep is None else self._g_keep.data_ptr(), self.dropout_p, self.bucket.data_ptr(),
                self._g_logits.data_ptr(), self._g_scratch.data_ptr(), st))
            allreduce_sum_(self.bucket, self.group)
            _lib.check(lib.vlmfw_probe_adamw_dev(
                self.params.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.bucket.data_ptr(),
                self.n_params, self.n_params + 1, self._g_steps_done.data_ptr(), self.lr, self.betas[0], self.betas[1],
                self.eps, self.wd, st))
            _lib.check(lib.vlmfw_probe_bn_running_update(
                self.running_mean.data_ptr(), self.running_var.data_ptr(), self.stats.data_ptr(), d, n_global,
                self.bn_momentum, st))
            _lib.check(lib.vlmfw_counter_add(self._g_steps_done.data_ptr(), 1, st))
            self._g_loss.copy_(self.bucket[self.n_params:], non_blocking=True)
        self._graph = graph

    def _capture_p2p(self, b_local: int, n_global: int, global_offset: int) -> None:
        """Graph form of the step with the exchange fused into its kernels (vlmfw_probe_linear_step_dp): seven launches of
        this library and nothing else. Communicator set-up: every rank allocates its exchange buffer, the cudaIpc handles go
        round with one all_gather_object, every rank maps its peers' buffers."""
        import ctypes as C

        import torch.distributed as dist

        dev, d, c, lib = self.device, self.dim, self.classes, self._lib
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        with torch.cuda.device(dev):
            h = C.c_void_p()
            _lib.check(lib.vlmfw_comm_create(rank, world, 2 * d, self.n_params + 2, C.byref(h)))
            self._comm_h = h
            if world > 1:
                nbytes = lib.vlmfw_comm_handle_bytes()
                mine = (C.c_char * nbytes)()
                _lib.check(lib.vlmfw_comm_handle(h, mine))
                handles = [None] * world
                dist.all_gather_object(handles, bytes(mine), group=self.group)
                blob = b"".join(handles)
                _lib.check(lib.vlmfw_comm_connect(h, blob))
                dist.barrier(self.group)  # every rank has mapped every buffer before anybody stores into one
        a = _lib.ProbeStep(
            x=self._g_x.data_ptr(), targets=self._g_y.data_ptr(), batch=b_local, dim=d, classes=c, n_global=n_global,
            row0=global_offset, stats=self.stats.data_ptr(), params=self.params.data_ptr(), exp_avg=self.exp_avg.data_ptr(),
            exp_avg_sq=self.exp_avg_sq.data_ptr(), bucket=self.bucket.data_ptr(), logits=self._g_logits.data_ptr(),
            scratch=self._g_scratch.data_ptr(), keep=None if self._g_keep is None else self._g_keep.data_ptr(),
            class_weight=None if self.class_weight is None else self.class_weight.data_ptr(),
            running_mean=self.running_mean.data_ptr(), running_var=self.running_var.data_ptr(),
            steps_done=self._g_steps_done.data_ptr(), mask_key=mask_key(self.seed, 0), mask_step_stride=MASK_KEY_STEP_STRIDE,
            dropout_p=float(self.dropout_p), bn_eps=float(self.bn_eps), bn_momentum=float(self.bn_momentum), lr=float(self.lr),
            beta1=float(self.betas[0]), beta2=float(self.betas[1]), eps=float(self.eps), weight_decay=float(self.wd))
        self._g_args = a
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.cuda.graph(graph):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.vlmfw_probe_linear_step_dp(C.byref(a), self._comm_h, st))
            self._g_loss.copy_(self.bucket[self.n_params:], non_blocking=True)
        self._graph = graph

    def comm_error(self) -> bool:
        """True when a wait of the fused exchange gave up (a peer stopped reporting): results since then are invalid."""
        import ctypes as C

        if self._comm_h is None:
            return False
        flag = C.c_int32(0)
        _lib.check(self._lib.vlmfw_comm_error(self._comm_h, C.byref(flag)))
        return bool(flag.value)

    def release_graph(self) -> None:
        """Drop the captured graph (and with it the captured NCCL kernels / the peer-memory communicator). Call before
        ``torch.distributed.destroy_process_group()``: tearing a communicator down under a live graph can block. In a
        process group every rank must call it (the peer buffers are unmapped behind a barrier)."""
        import torch.distributed as dist

        self._graph = None
        if self._comm_h is not None:
            torch.cuda.synchronize(self.device)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.barrier(self.group)  # nobody frees a buffer a peer may still be storing into
            self._lib.vlmfw_comm_destroy(self._comm_h)
            self._comm_h = None

    def step_graph(self, x: torch.Tensor, y: torch.Tensor) -> float:
        """``step`` through the captured graph: stage this rank's rows (host or device tensors) into the graph's input
        buffers, replay, wait for the loss. Shapes are the ones given to ``capture``."""
        if self._graph is None:
            raise RuntimeError("step_graph before capture()")
        if tuple(x.shape) != (self._g_shape[0], self.dim) or y.shape[0] != self._g_shape[0]:
            raise ValueError(f"step_graph: batch {tuple(x.shape)} does not match the captured shape {self._g_shape[0]}")
        if self._g_steps_done_value != self.step_count:  # eager steps ran in between
            self._g_steps_done.fill_(self.step_count)
        self._g_x.copy_(x, non_blocking=True)
        self._g_y.copy_(y, non_blocking=True)
        self._graph.replay()
        self.last_logits = self._g_logits
        self.step_count += 1
        self._g_steps_done_value = self.step_count
        self.num_batches_tracked += 1
        torch.cuda.current_stream(self.device).synchronize()
        loss_sum, weight_sum = self._g_loss.tolist()
        return loss_sum / weight_sum if weight_sum > 0 else 0.0

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


class _HeadState:
    """Flat fp32 parameters / AdamW moments / BatchNorm running statistics of one head.
    linear: [W (C*D) | b (C) | gamma (D) | beta (D)]; deeper: [W1 (Hd*D) | b1 (Hd) | W2 (C*Hd) | b2 (C) | gamma | beta]."""

    def __init__(self, dim: int, classes: int, hidden: int, device, gen: torch.Generator):
        self.dim, self.classes, self.hidden = dim, classes, hidden
        if hidden:
            self.segments = [("2.weight", hidden * dim, (hidden, dim)), ("2.bias", hidden, (hidden,)),
                             ("5.weight", classes * hidden, (classes, hidden)), ("5.bias", classes, (classes,))]
        else:
            self.segments = [("2.weight", classes * dim, (classes, dim)), ("2.bias", classes, (classes,))]
        self.segments += [("0.weight", dim, (dim,)), ("0.bias", dim, (dim,))]
        self.n_params = sum(n for _, n, _ in self.segments)
        self.params = torch.zeros(self.n_params, device=device)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.running_mean = torch.zeros(dim, device=device)
        self.running_var = torch.ones(dim, device=device)
        self.num_batches_tracked = 0
        # nn.Linear default init (uniform +-1/sqrt(fan_in)), BatchNorm1d weight 1 / bias 0, seeded
        for name, _, shape in self.segments:
            v = self.view(name)
            if name == "0.weight":
                v.fill_(1.0)
            elif name != "0.bias":
                fan_in = dim if name.startswith("2.") else hidden
                v.copy_((torch.rand(shape, generator=gen) * 2 - 1) / fan_in ** 0.5)

    def view(self, name: str) -> torch.Tensor:
        off = 0
        for key, n, shape in self.segments:
            if key == name:
                return self.params[off: off + n].view(shape)
            off += n
        raise KeyError(name)

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        for name, _, _ in self.segments:
            self.view(name).copy_(sd[prefix + name])
        self.running_mean.copy_(sd[prefix + "0.running_mean"])
        self.running_var.copy_(sd[prefix + "0.running_var"])
        self.num_batches_tracked = int(sd.get(prefix + "0.num_batches_tracked", 0))

    def state_dict(self, prefix: str = "") -> Dict[str, torch.Tensor]:
        out = {prefix + name: self.view(name).clone() for name, _, _ in self.segments}
        out[prefix + "0.running_mean"] = self.running_mean.clone()
        out[prefix + "0.running_var"] = self.running_var.clone()
        out[prefix + "0.num_batches_tracked"] = torch.tensor(self.num_batches_tracked)
        return out


class MultiTaskProbeTrainStep:
    """Train step of a ``MultiTaskProbe`` (or a deeper ``LinearProbe``) on cached features, data parallel.

    Restates the loop body of probing/train/multitask_trainer.py:188-263 over probing/train/base_trainer.py:166-181:
    every head sees the same features, its cross entropy runs over the rows whose label is present (``-1`` = missing),
    class-weighted (:162), the total loss is ``sum_t w_t * loss_t`` with static task weights or the inverse EMA of the
    per-task losses normalised to mean 1 (``_compute_task_weights``, :209-226; recomputed by ``on_epoch_start()``, the EMA
    fed after every step, :243-263), one AdamW over all head parameters. Heads: ``make_head`` or ``make_head_deeper``
    (probing/models/base_probe.py:7-22). Phases as in ``ProbeTrainStep``: shared BatchNorm sums -> all-reduce ->
    per-head gradients into ONE flat bucket -> all-reduce -> AdamW per head with ``grad = w_t * bucket / weight_sum_t``.
    State-dict keys are the reference's (``heads.<task>.{0,2[,5]}.*``).

    Two deliberate deviations, both only when a task has NO labelled row in a whole global batch: the reference leaves
    that head's parameters without gradient (skipped by AdamW, its step counter not advanced); here the update is
    skipped too but the bias-correction step counter is shared by all heads.
    """

    def __init__(self, dim: int, tasks: Dict[str, int], device="cuda", deeper: bool = False, hidden_dim: int = 512,
                 lr: float = 1e-4, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 dropout_p: float = 0.3, class_weights: Optional[Dict[str, torch.Tensor]] = None,
                 task_weights: Optional[Dict[str, float]] = None, use_running_means: bool = False,
                 ema_alpha: float = 0.99, bn_eps: float = 1e-5, bn_momentum: float = 0.1, seed: int = 0, group=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MultiTaskProbeTrainStep runs on CUDA only: vlmfw_b200 has no CPU fallback")
        self.dim, self.tasks = dim, dict(tasks)
        self.hidden = int(hidden_dim) if deeper else 0
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.dropout_p, self.bn_eps, self.bn_momentum = dropout_p, bn_eps, bn_momentum
        self.seed, self.group = seed, group
        self._lib = _lib.load()
        g = torch.Generator().manual_seed(seed)
        self.heads = {t: _HeadState(dim, c, self.hidden, self.device, g) for t, c in self.tasks.items()}
        self.class_weights = {t: (None if not class_weights or class_weights.get(t) is None
                                  else class_weights[t].to(self.device, torch.float32).contiguous()) for t in self.tasks}
        self.static_task_weights = {t: float((task_weights or {}).get(t, 1.0)) for t in self.tasks}
        self.use_running_means, self.ema_alpha = use_running_means, ema_alpha
        self.ema: Dict[str, Optional[float]] = {t: None for t in self.tasks}
        self.current_task_weights = dict(self.static_task_weights)
        self.stats = torch.zeros(2 * dim, dtype=torch.float64, device=self.device)
        sizes = [h.n_params + 2 for h in self.heads.values()]
        self.bucket = torch.zeros(sum(sizes), device=self.device)
        self._offsets = [sum(sizes[:i]) for i in range(len(sizes))]
        self.step_count = 0
        self.last_logits: Dict[str, torch.Tensor] = {}
        self._graph = None  # set by capture()

    # ---- task weights (multitask_trainer.py:209-226, 235-263) ---------------------------------------------------
    def on_epoch_start(self) -> Dict[str, float]:
        if not self.use_running_means:
            self.current_task_weights = dict(self.static_task_weights)
        else:
            raw = [1.0 / max(self.static_task_weights[t] if self.ema[t] is None else float(self.ema[t]), 1e-8)
                   for t in self.tasks]
            avg = sum(raw) / max(1, len(raw))
            self.current_task_weights = {t: raw[i] / avg for i, t in enumerate(self.tasks)}
        return dict(self.current_task_weights)

    def _update_ema(self, losses: Dict[str, Optional[float]]) -> None:
        for t, v in losses.items():
            if v is None or v != v or v in (float("inf"), float("-inf")):
                continue
            self.ema[t] = v if self.ema[t] is None else self.ema_alpha * self.ema[t] + (1 - self.ema_alpha) * v

    # ---- state --------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "heads.") -> None:
        for t, h in self.heads.items():
            h.load_state_dict(sd, f"{prefix}{t}.")

    def state_dict(self, prefix: str = "heads.") -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        for t, h in self.heads.items():
            out.update(h.state_dict(f"{prefix}{t}."))
        return out

    def dropout_keep_masks(self, task_index: int, n_global: int, lo: int, hi: int):
        """(keep1 [rows, D], keep2 [rows, Hd] or None), keyed by (seed, step, task, GLOBAL row)."""
        if self.dropout_p <= 0.0:
            return None, None
        key = mask_key(self.seed, self.step_count, task_index)
        k1 = device_keep_mask(self._lib, self.device, hi - lo, self.dim, lo, key, self.dropout_p)
        k2 = device_keep_mask(self._lib, self.device, hi - lo, self.hidden, lo, key + 1, self.dropout_p) if self.hidden else None
        return k1, k2

    # ---- phases -------------------------------------------------------------------------------------------------
    def phase1_stats(self, x: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(x.device):
            _lib.check(self._lib.vlmfw_probe_bn_stats(x.data_ptr(), x.shape[0], self.dim, self.stats.data_ptr(),
                                                      torch.cuda.current_stream(x.device).cuda_stream))
        return self.stats

    def phase2_grads(self, x: torch.Tensor, y: torch.Tensor, n_global: int, masks) -> torch.Tensor:
        """y int64 [B, T] (task order of ``self.tasks``); masks: list of (keep1, keep2) per task."""
        b = x.shape[0]
        st = torch.cuda.current_stream(x.device).cuda_stream
        yt = y.t().contiguous()  # [T, B]
        keepalive = [yt]
        with torch.cuda.device(x.device):
            for i, (t, h) in enumerate(self.heads.items()):
                bucket = self.bucket[self._offsets[i]: self._offsets[i] + h.n_params + 2]
                logits = torch.empty(b, h.classes, device=x.device)
                self.last_logits[t] = logits
                k1, k2 = masks[i]
                cw = self.class_weights[t]
                if self.hidden:
                    n = int(self._lib.vlmfw_probe_deep_scratch_floats(b, self.dim, self.hidden, h.classes))
                    scratch = torch.empty(n, device=x.device)
                    _lib.check(self._lib.vlmfw_probe_deep_grads(
                        x.data_ptr(), yt[i].data_ptr(), b, self.dim, self.hidden, h.classes, self.stats.data_ptr(),
                        n_global, self.bn_eps, h.params.data_ptr(), None if cw is None else cw.data_ptr(),
                        None if k1 is None else k1.data_ptr(), None if k2 is None else k2.data_ptr(), self.dropout_p,
                        bucket.data_ptr(), logits.data_ptr(), scratch.data_ptr(), st))
                else:
                    scratch = torch.empty(b * (h.classes + 2), device=x.device)
                    _lib.check(self._lib.vlmfw_probe_linear_grads(
                        x.data_ptr(), yt[i].data_ptr(), b, self.dim, h.classes, self.stats.data_ptr(), n_global,
                        self.bn_eps, h.view("0.weight").data_ptr(), h.view("0.bias").data_ptr(),
                        h.view("2.weight").data_ptr(), h.view("2.bias").data_ptr(),
                        None if cw is None else cw.data_ptr(), None if k1 is None else k1.data_ptr(), self.dropout_p,
                        bucket.data_ptr(), logits.data_ptr(), scratch.data_ptr(), st))
                keepalive += [scratch, k1, k2]
        self._keepalive = keepalive
        return self.bucket

    def phase3_update(self, n_global: int) -> Dict[str, Optional[float]]:
        self.step_count += 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            for i, (t, h) in enumerate(self.heads.items()):
                bucket = self.bucket[self._offsets[i]: self._offsets[i] + h.n_params + 2]
                _lib.check(self._lib.vlmfw_probe_adamw_ex(
                    h.params.data_ptr(), h.exp_avg.data_ptr(), h.exp_avg_sq.data_ptr(), bucket.data_ptr(), h.n_params,
                    h.n_params + 1, float(self.current_task_weights[t]), self.step_count, self.lr, self.betas[0],
                    self.betas[1], self.eps, self.wd, st))
                _lib.check(self._lib.vlmfw_probe_bn_running_update(
                    h.running_mean.data_ptr(), h.running_var.data_ptr(), self.stats.data_ptr(), self.dim, n_global,
                    self.bn_momentum, st))
                h.num_batches_tracked += 1
        # the step's only D2H read: [loss_sum, weight_sum] of every head
        idx = torch.tensor([o + h.n_params + k for o, h in zip(self._offsets, self.heads.values()) for k in (0, 1)],
                           device=self.device)
        tails = self.bucket[idx].tolist()
        losses = {t: (tails[2 * i] / tails[2 * i + 1] if tails[2 * i + 1] > 0 else None) for i, t in enumerate(self.tasks)}
        if self.use_running_means:
            self._update_ema(losses)
        return losses

    # ---- the whole step as ONE CUDA graph -----------------------------------------------------------------------
    # launches of this library per graph-form step: bn_stats, then per head 2 (linear: 1) mask draws + the gradient kernels
    # (deeper: 13, linear: 2), then per head AdamW + BatchNorm running statistics, the step counter
    def kernels_per_step(self) -> int:
        per_head = (2 + 13 if self.hidden else 1 + 2) if self.dropout_p > 0 else (13 if self.hidden else 2)
        return 1 + len(self.tasks) * (per_head + 2) + 1

    def capture(self, b_local: int, n_global: Optional[int] = None, global_offset: int = 0) -> None:
        """Capture one optimisation step (all heads) for a fixed local batch size into a CUDA graph, like
        ``ProbeTrainStep.capture``: the dropout keys and AdamW's bias corrections come from a device-side step counter
        (``vlmfw_dropout_keep_mask_dev``, ``vlmfw_probe_adamw_ex_dev``), the task weights from a device vector that
        ``on_epoch_start`` refreshes, the per-head [loss_sum, weight_sum] pairs land in pinned memory by a copy node. A
        5-head deeper step is ~80 launches eagerly; replayed it is one."""
        import torch.distributed as dist

        dev, d, lib = self.device, self.dim, self._lib
        n_global = b_local if n_global is None else n_global
        t_count = len(self.tasks)
        self._g_shape = (b_local, n_global, global_offset)
        self._g_x = torch.zeros(b_local, d, device=dev)
        self._g_y = torch.full((t_count, b_local), -1, dtype=torch.int64, device=dev)  # task-major, like phase2's y.t()
        self._g_steps_done = torch.tensor([self.step_count], dtype=torch.int64, device=dev)
        self._g_steps_done_value = self.step_count
        self._g_tw_value = [float(self.current_task_weights[t]) for t in self.tasks]
        self._g_tw = torch.tensor(self._g_tw_value, dtype=torch.float32, device=dev)
        self._g_tails_idx = torch.tensor([o + h.n_params + k for o, h in zip(self._offsets, self.heads.values())
                                          for k in (0, 1)], device=dev)
        self._g_tails_dev = torch.zeros(2 * t_count, device=dev)
        self._g_tails = torch.zeros(2 * t_count, dtype=torch.float32).pin_memory()
        bufs = []
        for h in self.heads.values():
            n = int(lib.vlmfw_probe_deep_scratch_floats(b_local, d, self.hidden, h.classes)) if self.hidden \
                else b_local * (h.classes + 2)
            bufs.append({
                "logits": torch.empty(b_local, h.classes, device=dev), "scratch": torch.empty(n, device=dev),
                "k1": torch.empty(b_local, d, dtype=torch.uint8, device=dev) if self.dropout_p > 0 else None,
                "k2": torch.empty(b_local, self.hidden, dtype=torch.uint8, device=dev)
                if (self.dropout_p > 0 and self.hidden) else None})
        self._g_bufs = bufs
        self._g_streams = [torch.cuda.Stream(device=dev) for _ in self.heads]
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if distributed:  # communicator set-up cannot be captured
            allreduce_sum_(torch.zeros(2, dtype=torch.float64, device=dev), self.group)
            allreduce_sum_(torch.zeros(2, device=dev), self.group)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.cuda.graph(graph):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.vlmfw_probe_bn_stats(self._g_x.data_ptr(), b_local, d, self.stats.data_ptr(), st))
            allreduce_sum_(self.stats, self.group)
            # the heads are independent between the two exchanges: each head's chain (mask draws, forward, backward — 15
            # launches of small grids for a deeper head) is captured on its own stream, so the graph holds one branch per
            # head and the GPU runs them side by side (a 5-head deeper step: 1.75 ms as one serial chain)
            main = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            for i, (t, h) in enumerate(self.heads.items()):
                branch = self._g_streams[i]
                branch.wait_event(fork)
                with torch.cuda.stream(branch):
                    bst = branch.cuda_stream
                    bf = bufs[i]
                    bucket = self.bucket[self._offsets[i]: self._offsets[i] + h.n_params + 2]
                    cw = self.class_weights[t]
                    key0 = mask_key(self.seed, 0, i)
                    if bf["k1"] is not None:
                        _lib.check(lib.vlmfw_dropout_keep_mask_dev(bf["k1"].data_ptr(), b_local, d, global_offset, key0,
                                                                   self._g_steps_done.data_ptr(), MASK_KEY_STEP_STRIDE,
                                                                   float(self.dropout_p), bst))
                    if bf["k2"] is not None:
                        _lib.check(lib.vlmfw_dropout_keep_mask_dev(bf["k2"].data_ptr(), b_local, self.hidden, global_offset,
                                                                   (key0 + 1) & 0xFFFFFFFFFFFFFFFF,
                                                                   self._g_steps_done.data_ptr(), MASK_KEY_STEP_STRIDE,
                                                                   float(self.dropout_p), bst))
                    k1 = None if bf["k1"] is None else bf["k1"].data_ptr()
                    k2 = None if bf["k2"] is None else bf["k2"].data_ptr()
                    if self.hidden:
                        _lib.check(lib.vlmfw_probe_deep_grads(
                            self._g_x.data_ptr(), self._g_y[i].data_ptr(), b_local, d, self.hidden, h.classes,
                            self.stats.data_ptr(), n_global, self.bn_eps, h.params.data_ptr(),
                            None if cw is None else cw.data_ptr(), k1, k2, self.dropout_p, bucket.data_ptr(),
                            bf["logits"].data_ptr(), bf["scratch"].data_ptr(), bst))
                    else:
                        _lib.check(lib.vlmfw_probe_linear_grads(
                            self._g_x.data_ptr(), self._g_y[i].data_ptr(), b_local, d, h.classes, self.stats.data_ptr(), n_global,
                            self.bn_eps, h.view("0.weight").data_ptr(), h.view("0.bias").data_ptr(),
                            h.view("2.weight").data_ptr(), h.view("2.bias").data_ptr(), None if cw is None else cw.data_ptr(),
                            k1, self.dropout_p, bucket.data_ptr(), bf["logits"].data_ptr(), bf["scratch"].data_ptr(), bst))
                    join = torch.cuda.Event()
                    join.record(branch)
                main.wait_event(join)
            allreduce_sum_(self.bucket, self.group)
            for i, (t, h) in enumerate(self.heads.items()):
                bucket = self.bucket[self._offsets[i]: self._offsets[i] + h.n_params + 2]
                _lib.check(lib.vlmfw_probe_adamw_ex_dev(
                    h.params.data_ptr(), h.exp_avg.data_ptr(), h.exp_avg_sq.data_ptr(), bucket.data_ptr(), h.n_params,
                    h.n_params + 1, self._g_tw.data_ptr() + 4 * i, self._g_steps_done.data_ptr(), self.lr, self.betas[0],
                    self.betas[1], self.eps, self.wd, st))
                _lib.check(lib.vlmfw_probe_bn_running_update(
                    h.running_mean.data_ptr(), h.running_var.data_ptr(), self.stats.data_ptr(), d, n_global,
                    self.bn_momentum, st))
            _lib.check(lib.vlmfw_counter_add(self._g_steps_done.data_ptr(), 1, st))
            torch.index_select(self.bucket, 0, self._g_tails_idx, out=self._g_tails_dev)
            self._g_tails.copy_(self._g_tails_dev, non_blocking=True)
        self._graph = graph

    def release_graph(self) -> None:
        """Drop the captured graph (call before ``torch.distributed.destroy_process_group()``)."""
        self._graph = None

    def step_graph(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, Optional[float]]:
        """``step`` through the captured graph: ``x [B_local, D]``, ``y [B_local, T]`` (host or device), shapes as captured."""
        if self._graph is None:
            raise RuntimeError("step_graph before capture()")
        if y.dim() == 1:
            y = y[:, None]
        if tuple(x.shape) != (self._g_shape[0], self.dim) or tuple(y.shape) != (self._g_shape[0], len(self.tasks)):
            raise ValueError(f"step_graph: batch {tuple(x.shape)} / {tuple(y.shape)} does not match the captured shape")
        if self._g_steps_done_value != self.step_count:  # eager steps ran in between
            self._g_steps_done.fill_(self.step_count)
        tw = [float(self.current_task_weights[t]) for t in self.tasks]
        if tw != self._g_tw_value:
            self._g_tw.copy_(torch.tensor(tw, dtype=torch.float32))
            self._g_tw_value = tw
        self._g_x.copy_(x, non_blocking=True)
        self._g_y.copy_(y.t(), non_blocking=True)
        self._graph.replay()
        self.last_logits = {t: self._g_bufs[i]["logits"] for i, t in enumerate(self.tasks)}
        self.step_count += 1
        self._g_steps_done_value = self.step_count
        for h in self.heads.values():
            h.num_batches_tracked += 1
        torch.cuda.current_stream(self.device).synchronize()
        tails = self._g_tails.tolist()
        losses = {t: (tails[2 * i] / tails[2 * i + 1] if tails[2 * i + 1] > 0 else None) for i, t in enumerate(self.tasks)}
        if self.use_running_means:
            self._update_ema(losses)
        return losses

    def step(self, x: torch.Tensor, y: torch.Tensor, n_global: Optional[int] = None, global_offset: int = 0,
             masks=None) -> Dict[str, Optional[float]]:
        """One optimisation step on this rank's rows: ``x [B_local, D]`` fp32, ``y [B_local, T]`` int64 (-1 = missing).
        Returns the per-task mean losses over the GLOBAL batch (None where no row carried the label)."""
        if not x.is_cuda:
            raise RuntimeError("features must be CUDA tensors")
        x = x.to(torch.float32).contiguous()
        y = y.to(device=x.device, dtype=torch.int64)
        if y.dim() == 1:
            y = y[:, None]
        b = x.shape[0]
        n_global = b if n_global is None else n_global
        if masks is None:
            masks = [self.dropout_keep_masks(i, n_global, global_offset, global_offset + b) for i in range(len(self.tasks))]
        allreduce_sum_(self.phase1_stats(x), self.group)
        allreduce_sum_(self.phase2_grads(x, y, n_global, masks), self.group)
        return self.phase3_update(n_global)
