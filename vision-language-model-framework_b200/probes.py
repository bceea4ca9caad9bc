"""Probe heads on pooled features — same classes, constructor arguments and state-dict keys as the reference
(probing/models/base_probe.py:7-83, linear_probe.py:8-48, multitask_probe.py:7-48), with the eval-mode head arithmetic
(BatchNorm1d(eval) -> Linear [-> GELU -> Linear] -> argmax, every head in ONE launch) running in
``vlmfw_probe_forward``.

Training through these modules (autograd over the tiny heads) stays plain ``nn.Sequential`` so the reference trainers
keep working unchanged; the native data-parallel train step on cached features is ``probe_train.ProbeTrainStep``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .vision_backbone import VisionBackbone


def make_head(in_dim: int, out_dim: int, dropout_p: float) -> nn.Sequential:
    return nn.Sequential(nn.BatchNorm1d(in_dim), nn.Dropout(p=dropout_p), nn.Linear(in_dim, out_dim))


def make_head_deeper(in_dim: int, out_dim: int, hidden_dim: int, dropout_p: float) -> nn.Sequential:
    return nn.Sequential(nn.BatchNorm1d(in_dim), nn.Dropout(p=dropout_p), nn.Linear(in_dim, hidden_dim), nn.GELU(),
                         nn.Dropout(p=dropout_p), nn.Linear(hidden_dim, out_dim))


def _head_desc(head: nn.Sequential, logit_offset: int) -> Tuple[_lib.Head, List[torch.Tensor]]:
    bn, lin1 = head[0], head[2]
    deeper = len(head) == 6
    lin2 = head[5] if deeper else None
    keep = [bn.weight, bn.bias, bn.running_mean, bn.running_var, lin1.weight, lin1.bias]
    if deeper:
        keep += [lin2.weight, lin2.bias]
    for t in keep:
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise RuntimeError("probe heads must be contiguous fp32 CUDA parameters for the fused path")
    d = _lib.Head(
        bn_weight=bn.weight.data_ptr(), bn_bias=bn.bias.data_ptr(), bn_mean=bn.running_mean.data_ptr(),
        bn_var=bn.running_var.data_ptr(), w1=lin1.weight.data_ptr(), b1=lin1.bias.data_ptr(),
        w2=lin2.weight.data_ptr() if deeper else None, b2=lin2.bias.data_ptr() if deeper else None,
        classes=(lin2 if deeper else lin1).out_features, hidden=lin1.out_features if deeper else 0, bn_eps=bn.eps,
        logit_offset=logit_offset)
    return d, keep


def fused_heads_forward(feats: torch.Tensor, heads: Sequence[nn.Sequential],
                        want_argmax: bool = True) -> Tuple[List[torch.Tensor], Optional[torch.Tensor]]:
    """feats fp32 [B, D] (CUDA) -> ([logits_h [B, C_h] ...], argmax int32 [B, n_heads]) in one kernel launch."""
    if not feats.is_cuda:
        raise RuntimeError("fused_heads_forward needs CUDA features: vlmfw_b200 has no CPU fallback")
    feats = feats.to(torch.float32).contiguous()
    b, dim = feats.shape
    descs = (_lib.Head * len(heads))()
    alive, off, widths = [], 0, []
    for i, h in enumerate(heads):
        descs[i], keep = _head_desc(h, off)
        alive.append(keep)
        widths.append(descs[i].classes)
        off += descs[i].classes
    logits = torch.empty(b, off, dtype=torch.float32, device=feats.device)
    amax = torch.empty(b, len(heads), dtype=torch.int32, device=feats.device) if want_argmax else None
    if b > 0:
        with torch.cuda.device(feats.device):
            _lib.check(_lib.load().vlmfw_probe_forward(
                feats.data_ptr(), b, dim, descs, len(heads), logits.data_ptr(), off,
                None if amax is None else amax.data_ptr(), torch.cuda.current_stream(feats.device).cuda_stream))
    parts = list(logits.split(widths, dim=1))
    if len(parts) > 1:  # the reference's heads return contiguous [B, C_t] tensors
        parts = [p.contiguous() for p in parts]
    return parts, amax


class BaseProbe(nn.Module):
    """Backbone + freeze bookkeeping (reference probing/models/base_probe.py:25-83)."""

    def __init__(self, backbone: VisionBackbone, freeze_backbone: bool = True):
        super().__init__()
        self.backbone = backbone
        self.set_freeze_backbone(bool(freeze_backbone))
        self._fully_frozen = self._compute_fully_frozen()

    def _compute_fully_frozen(self) -> bool:
        return not any(p.requires_grad for p in self.backbone.parameters())

    @staticmethod
    def _align_dtype(x: torch.Tensor, ref_module: nn.Module) -> torch.Tensor:
        try:
            target_dtype = next(ref_module.parameters()).dtype
        except StopIteration:
            return x
        return x if x.dtype == target_dtype else x.to(target_dtype)

    def extract_features(self, images):
        if self._fully_frozen:
            was_training = self.backbone.training
            self.backbone.eval()
            with torch.no_grad():
                feats = self.backbone(images)
            if was_training:
                self.backbone.train()
            return feats
        return self.backbone(images)

    def unfreeze_last_backbone_k_layers(self, k: int, parts: str = "all", include_embeddings: bool = True):
        self.backbone.unfreeze_last_k_layers(k=k, parts=parts, include_embeddings=include_embeddings)
        self._fully_frozen = self._compute_fully_frozen()

    def set_freeze_backbone(self, freeze: bool):
        for p in self.backbone.parameters():
            p.requires_grad = not freeze
        self._fully_frozen = self._compute_fully_frozen()

    @torch.no_grad()
    def predict_stream(self, batches):
        """``predict`` over an iterable of batches as a generator, software-pipelined so the GPU never waits for the host:
        the host->device copy of batch i+1 runs on the backbone's copy stream while batch i computes
        (``B200VisionBackbone.stage_async``), and the labels of batch i are read back (pinned buffer, async copy) after
        batch i+1 has been enqueued. Yields, in order, exactly what ``predict(batch)`` returns, as CPU tensors (or a dict
        of CPU tensors for a multitask probe). The reference's tester loop (probing/test/base_tester.py:47-61) is the
        serial form of this: one blocking ``.cpu()`` per batch (singletask_tester.py:98)."""
        stage = getattr(self.backbone, "stage_async", None)
        it = iter(batches)

        def staged():
            try:
                nxt = next(it)
            except StopIteration:
                return
            cur = stage(nxt) if stage is not None else nxt
            while True:
                try:
                    nxt = next(it)
                except StopIteration:
                    yield cur
                    return
                ahead = stage(nxt) if stage is not None else nxt  # issued before `cur` is consumed: overlaps its forward
                yield cur
                cur = ahead

        ring, pending, k = [None, None], None, 0

        def finish(p):
            host, ev, names = p
            ev.synchronize()
            out = host.clone()
            return out if names is None else {t: out[:, i] for i, t in enumerate(names)}

        for batch in staged():
            pred = self.predict(batch)
            names = None
            if isinstance(pred, dict):
                names = list(pred.keys())
                pred = torch.stack([pred[t] for t in names], dim=1)
            slot = k & 1
            k += 1
            if ring[slot] is None or ring[slot].shape != pred.shape or ring[slot].dtype != pred.dtype:
                ring[slot] = torch.empty(pred.shape, dtype=pred.dtype).pin_memory() if pred.is_cuda else None
            if pred.is_cuda:
                ring[slot].copy_(pred, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(pred.device))
                cur = (ring[slot], ev, names)
            else:
                cur = None
            if pending is not None:
                yield finish(pending)
            if cur is None:  # CPU predictions (no device work to pipeline)
                yield pred if names is None else {t: pred[:, i] for i, t in enumerate(names)}
                pending = None
            else:
                pending = cur
        if pending is not None:
            yield finish(pending)

    # fused eval path is taken when nothing needs autograd and everything is fp32 on the GPU
    def _use_fused(self, feats: torch.Tensor, heads: Sequence[nn.Module]) -> bool:
        if not feats.is_cuda or torch.is_grad_enabled() and any(p.requires_grad for h in heads for p in h.parameters()):
            return False
        if any(h.training for h in heads):
            return False
        # limits of vlmfw_probe_forward (csrc/probe.cu): <= 16 heads, <= 32 classes per head, and the staged row
        # (dim + hidden + 32 floats) inside 48 KB of shared memory; anything else runs the nn.Sequential path
        if len(heads) > 16:
            return False
        for h in heads:
            deeper = len(h) == 6
            classes = (h[5] if deeper else h[2]).out_features
            hidden = h[2].out_features if deeper else 0
            if classes > 32 or (feats.shape[-1] + hidden + 32) * 4 > 48 * 1024:
                return False
        return all(p.dtype == torch.float32 and p.is_cuda for h in heads for p in h.parameters())


class LinearProbe(BaseProbe):
    """backbone + one head; ``forward(images) -> logits [B, C]`` (reference probing/models/linear_probe.py:8-48)."""

    def __init__(self, backbone: VisionBackbone, n_out_classes: int, freeze_backbone: bool = True,
                 dropout_p: float = 0.3, deeper_head: bool = False, hidden_dim: int = 512):
        super().__init__(backbone, freeze_backbone)
        if deeper_head:
            self.classifier = make_head_deeper(self.backbone.output_dim, n_out_classes, hidden_dim, dropout_p)
        else:
            self.classifier = make_head(self.backbone.output_dim, n_out_classes, dropout_p)

    def forward(self, images, **kwargs):
        feats = self._align_dtype(self.extract_features(images), self.classifier)
        if self._use_fused(feats, [self.classifier]):
            return fused_heads_forward(feats, [self.classifier], want_argmax=False)[0][0]
        return self.classifier(feats)

    @torch.no_grad()
    def predict(self, images):
        self.eval()
        feats = self._align_dtype(self.extract_features(images), self.classifier)
        if self._use_fused(feats, [self.classifier]):
            return fused_heads_forward(feats, [self.classifier])[1][:, 0].to(torch.int64)
        return torch.argmax(self.classifier(feats), dim=-1)


class MultiTaskProbe(BaseProbe):
    """backbone + one head per task; ``forward -> {"logits": {task: [B, C_t]}}``
    (reference probing/models/multitask_probe.py:7-48)."""

    def __init__(self, backbone: VisionBackbone, tasks: Dict[str, int], freeze_backbone: bool = True,
                 dropout_p: float = 0.3, deeper_heads: bool = False, hidden_dim: int = 512):
        super().__init__(backbone, freeze_backbone)
        self.tasks = dict(tasks)
        if deeper_heads:
            self.heads = nn.ModuleDict({t: make_head_deeper(self.backbone.output_dim, c, hidden_dim, dropout_p)
                                        for t, c in self.tasks.items()})
        else:
            self.heads = nn.ModuleDict({t: make_head(self.backbone.output_dim, c, dropout_p)
                                        for t, c in self.tasks.items()})

    def _run(self, images, want_argmax: bool):
        feats = self._align_dtype(self.extract_features(images), self.heads)
        mods = [self.heads[t] for t in self.tasks]
        if self._use_fused(feats, mods):
            logits, amax = fused_heads_forward(feats, mods, want_argmax=want_argmax)
            return dict(zip(self.tasks, logits)), amax
        return {t: self.heads[t](feats) for t in self.tasks}, None

    def forward(self, images, **kwargs):
        return {"logits": self._run(images, False)[0]}

    @torch.no_grad()
    def predict(self, images):
        self.eval()
        logits, amax = self._run(images, True)
        if amax is not None:
            return {t: amax[:, i].to(torch.int64) for i, t in enumerate(self.tasks)}
        return {t: lg.argmax(dim=1) for t, lg in logits.items()}


def parsed_attributes(pred, task_names: Optional[Sequence[str]] = None) -> List[Dict[str, int]]:
    """The ``[{task: int}, ...]`` list the testers emit (reference probing/test/base_tester.py:53-57,
    singletask_tester.py:95-98, multitask_tester.py:91-95): one D2H copy for the whole batch."""
    if isinstance(pred, dict):
        names = list(pred.keys())
        cols = torch.stack([pred[t] for t in names], dim=1).cpu().tolist()
        return [{t: int(v) for t, v in zip(names, row)} for row in cols]
    name = task_names[0] if task_names else "label"
    return [{name: int(v)} for v in pred.cpu().tolist()]
