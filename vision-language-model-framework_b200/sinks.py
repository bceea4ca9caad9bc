"""Consumers on either side of the hot path that the reference runs with one host sync per batch:

``PredictionSink``      the tester's inference loop body — probing/test/base_tester.py:47-61 (``pred_idxs[i]`` ->
                        ``preds.append({key: int(...)})``) and singletask_tester.py:95-98 (``.argmax(dim=1).cpu()
                        .tolist()`` every batch) — plus the metrics of datasets_vlm/evaluate_dataset.py:28-200
                        (accuracy, label set, average accuracy; confusion matrix counts). Labels stay on the GPU until
                        the dataset is done: one D2H copy per dataset, confusion counts by ``vlmfw_confusion_update``.
``FeatureCacheWriter``  the frozen-feature extraction loop of probing/train/singletask_trainer.py:219-285
                        (``feats_all.append(feats_b.detach().cpu())`` every batch, then ``torch.save({"x","y"})``):
                        features are staged in device chunks that drain to pinned host memory on a side stream, each
                        rank writes its shard, rank 0 merges them into the reference's file format.
``load_feature_cache``  the reading half (:233-242), same accepted keys.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

AGE_CLASS_NAMES = ["0-2", "3-9", "10-19", "20-29", "30-39", "40-49", "50-59", "60-69", "70+"]  # evaluate_dataset.py:7
FACE_ACCURACY_TASKS = ("gender", "ethnicity", "emotion")  # evaluate_dataset.py:126
MISSING_LABEL = -1  # probing/train/utils.py:8


def targets_to_matrix(targets_list: Sequence[Mapping], tasks: Sequence[str]) -> torch.Tensor:
    """List of per-sample target dicts -> int64 [B, T], -1 for missing (reference probing/train/utils.py:32-43)."""
    out = torch.empty(len(targets_list), len(tasks), dtype=torch.int64)
    for j, task in enumerate(tasks):
        for i, t in enumerate(targets_list):
            v = t.get(task, None)
            out[i, j] = int(v) if v is not None else MISSING_LABEL
    return out


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class PredictionSink:
    """Accumulates predicted labels (and ground truth) of a whole dataset on the device.

    ``add(pred, targets)``: ``pred`` is what ``LinearProbe.predict`` / ``MultiTaskProbe.predict`` return (int tensor [B],
    or ``{task: [B]}``), or the raw int32 ``[B, T]`` argmax of ``fused_heads_forward``; ``targets`` the batch's list of
    target dicts (or an int64 ``[B, T]`` tensor). Nothing synchronises until ``finalize()``.
    """

    def __init__(self, tasks: Sequence[str], classes: Union[int, Mapping[str, int]], device="cuda",
                 capacity: int = 4096, group=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PredictionSink keeps labels on the GPU: vlmfw_b200 has no CPU fallback")
        self.tasks = list(tasks)
        per_task = {t: int(classes) for t in self.tasks} if isinstance(classes, int) else {t: int(classes[t]) for t in self.tasks}
        self.classes = per_task
        self.cmax = max(per_task.values())
        self.group = group
        self._lib = _lib.load()
        t = len(self.tasks)
        self._pred = torch.empty(capacity, t, dtype=torch.int32, device=self.device)
        self._gt = torch.empty(capacity, t, dtype=torch.int64, device=self.device)
        self._n = 0
        self._counts = torch.zeros(t, self.cmax, self.cmax, dtype=torch.int32, device=self.device)
        self._dropped = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._final: Optional[dict] = None

    def _grow(self, need: int) -> None:
        if need <= self._pred.shape[0]:
            return
        cap = max(need, 2 * self._pred.shape[0])
        for name in ("_pred", "_gt"):
            old = getattr(self, name)
            new = torch.empty(cap, old.shape[1], dtype=old.dtype, device=self.device)
            new[: self._n] = old[: self._n]
            setattr(self, name, new)

    def add(self, pred, targets) -> None:
        if self._final is not None:
            raise RuntimeError("PredictionSink.add after finalize()")
        if isinstance(pred, Mapping):
            pred = torch.stack([pred[t] for t in self.tasks], dim=1)
        if pred.dim() == 1:
            pred = pred.unsqueeze(1)
        if not pred.is_cuda:
            raise RuntimeError("predictions must be CUDA tensors (the per-batch .cpu() is what this sink removes)")
        b = pred.shape[0]
        if pred.shape[1] != len(self.tasks):
            raise ValueError(f"pred has {pred.shape[1]} columns for {len(self.tasks)} tasks")
        gt = targets if isinstance(targets, torch.Tensor) else targets_to_matrix(targets, self.tasks)
        if gt.dim() == 1:
            gt = gt.unsqueeze(1)
        if tuple(gt.shape) != (b, len(self.tasks)):
            raise ValueError(f"targets shape {tuple(gt.shape)} does not match predictions {tuple(pred.shape)}")
        self._grow(self._n + b)
        ps = self._pred[self._n: self._n + b]
        gs = self._gt[self._n: self._n + b]
        ps.copy_(pred)  # D2D (+ int64 -> int32)
        if gt.device.type == "cpu":
            gt = gt.to(torch.int64).contiguous().pin_memory()
        gs.copy_(gt, non_blocking=True)
        if b:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.vlmfw_confusion_update(
                    ps.data_ptr(), gs.data_ptr(), b, len(self.tasks), self.cmax, self._counts.data_ptr(),
                    self._dropped.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        self._n += b

    # ------------------------------------------------------------------------------------------------ results
    def finalize(self) -> dict:
        """The dataset's single device->host read. With a process group: rows are concatenated in rank order (the
        contiguous sharding of ``parallel.shard_range``) and the counts summed, on every rank."""
        if self._final is not None:
            return self._final
        pred, gt = self._pred[: self._n], self._gt[: self._n]
        counts, dropped = self._counts.clone(), self._dropped.clone()
        rank, world = _world(self.group)
        if world > 1:
            sizes = torch.zeros(world, dtype=torch.int64, device=self.device)
            sizes[rank] = self._n
            dist.all_reduce(sizes, group=self.group)
            sizes = sizes.tolist()
            mx = max(sizes)
            both = torch.full((mx, 2 * len(self.tasks)), MISSING_LABEL, dtype=torch.int64, device=self.device)
            both[: self._n, : len(self.tasks)] = pred
            both[: self._n, len(self.tasks):] = gt
            outs = [torch.empty_like(both) for _ in range(world)]
            dist.all_gather(outs, both, group=self.group)
            full = torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)
            pred, gt = full[:, : len(self.tasks)], full[:, len(self.tasks):]
            dist.all_reduce(counts, group=self.group)
            dist.all_reduce(dropped, group=self.group)
        self._final = {"pred": pred.cpu().numpy().astype(np.int64), "gt": gt.cpu().numpy().astype(np.int64),
                       "counts": counts.cpu().numpy().astype(np.int64), "dropped": int(dropped.item())}
        return self._final

    def preds_and_gts(self) -> Tuple[List[Dict[str, int]], List[Dict[str, int]]]:
        """``preds`` / ``gts`` exactly as BaseTester.run_one builds them (base_tester.py:53-57)."""
        f = self.finalize()
        p, g = f["pred"].tolist(), f["gt"].tolist()
        preds = [{t: int(v) for t, v in zip(self.tasks, row)} for row in p]
        gts = [{t: int(v) for t, v in zip(self.tasks, row)} for row in g]
        return preds, gts

    def confusion_matrix(self, task: str) -> np.ndarray:
        """[C_task, C_task] int64, rows = ground truth; over the rows whose label is present."""
        c = self.classes[task]
        return self.finalize()["counts"][self.tasks.index(task), :c, :c]

    def metrics(self, dataset_name: str = "MiviaPar", age_mode: str = "classification") -> dict:
        """The dict Evaluator writes to metrics.json, computed from the device-accumulated confusion counts."""
        return metrics_from_counts(self.tasks, self.finalize()["counts"], dataset_name, age_mode)

    def save(self, out_dir, dataset_name: str = "MiviaPar", age_mode: str = "classification") -> dict:
        """preds.json / gts.json / metrics.json, formatted like Evaluator._save_json (indent=4). Rank 0 writes."""
        m = self.metrics(dataset_name, age_mode)
        if _world(self.group)[0] == 0:
            out = Path(out_dir)
            out.mkdir(parents=True, exist_ok=True)
            preds, gts = self.preds_and_gts()
            for name, data in (("preds.json", preds), ("gts.json", gts), ("metrics.json", m)):
                with open(out / name, "w") as fh:
                    json.dump(data, fh, indent=4)
        return m


def metrics_from_counts(tasks: Sequence[str], counts: np.ndarray, dataset_name: str = "MiviaPar",
                        age_mode: str = "classification") -> dict:
    """Evaluator's metrics.json content (evaluate_dataset.py:70-97 MiviaPar: every task of ``preds[0]`` in order;
    :118-200 face datasets: gender / ethnicity / emotion, then age) from confusion counts ``[T, C, C]`` (rows = ground
    truth over the samples whose label is present). accuracy = trace / total, ``labels`` = sorted classes that occur as
    truth or prediction — what ``accuracy_score`` and ``sorted(set(y_true + y_pred))`` give on the lists."""
    tasks = list(tasks)
    metrics: Dict[str, object] = {}
    accs: List[float] = []
    if dataset_name == "MiviaPar":
        order = tasks
    else:
        order = [t for t in FACE_ACCURACY_TASKS if t in tasks] + (["age"] if "age" in tasks else [])
    for task in order:
        cm = np.asarray(counts[tasks.index(task)])
        total = int(cm.sum())
        if total == 0:
            continue
        acc = float(np.trace(cm)) / float(total)
        present = np.nonzero((cm.sum(axis=0) + cm.sum(axis=1)) > 0)[0].tolist()
        if task == "age" and dataset_name != "MiviaPar":
            if age_mode.lower() != "classification":
                raise ValueError("the sink carries class indices; age regression metrics need raw values")
            metrics[task] = {"mode": "classification", "accuracy": acc, "labels": AGE_CLASS_NAMES}
        else:
            metrics[task] = {"accuracy": acc, "labels": [int(v) for v in present]}
        accs.append(acc)
    metrics["average_accuracy"] = sum(accs) / len(accs) if accs else None
    return metrics


class FeatureCacheWriter:
    """Streams pooled features ``[B, D]`` (device, fp32) and labels into ``{split}_features.pt`` without a per-batch
    host sync. Chunks of ``chunk_rows`` rows fill on the device; a full chunk drains to pinned host memory on a side
    stream while the tower computes the next batches."""

    def __init__(self, dim: int, device="cuda", chunk_rows: int = 8192, group=None, dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FeatureCacheWriter stages features on the GPU: vlmfw_b200 has no CPU fallback")
        self.dim, self.chunk_rows, self.group, self.dtype = dim, int(chunk_rows), group, dtype
        self._side = torch.cuda.Stream(self.device)
        self._bufs = [torch.empty(self.chunk_rows, dim, dtype=dtype, device=self.device) for _ in range(2)]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]  # buffer i may be overwritten after this event
        self._cur, self._fill = 0, 0
        self._host: List[torch.Tensor] = []
        self._ys: List[torch.Tensor] = []
        self._first_use = [True, True]

    def _drain(self, rows: int) -> None:
        buf = self._bufs[self._cur]
        host = torch.empty(rows, self.dim, dtype=self.dtype).pin_memory()
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._side):
            host.copy_(buf[:rows], non_blocking=True)
            self._free[self._cur].record(self._side)
        self._first_use[self._cur] = False
        self._host.append(host)
        self._cur ^= 1
        self._fill = 0
        if not self._first_use[self._cur]:
            torch.cuda.current_stream(self.device).wait_event(self._free[self._cur])

    def add(self, feats: torch.Tensor, y) -> None:
        if not feats.is_cuda:
            raise RuntimeError("features must be CUDA tensors (the per-batch .cpu() is what this writer removes)")
        feats = feats.detach()
        y = torch.as_tensor(y, dtype=torch.int64).reshape(-1).cpu()
        if feats.shape[0] != y.shape[0] or feats.shape[1] != self.dim:
            raise ValueError(f"features {tuple(feats.shape)} / labels {tuple(y.shape)} mismatch (dim {self.dim})")
        self._ys.append(y)
        done = 0
        while done < feats.shape[0]:
            take = min(feats.shape[0] - done, self.chunk_rows - self._fill)
            self._bufs[self._cur][self._fill: self._fill + take].copy_(feats[done: done + take])
            self._fill += take
            done += take
            if self._fill == self.chunk_rows:
                self._drain(self.chunk_rows)

    def finalize(self, path=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Drains, synchronises once and returns this rank's ``(x [N_local, D], y [N_local])`` on the host. With
        ``path``: writes the reference's cache file ``{"x": [N, D], "y": [N]}`` (singletask_trainer.py:279-283) — in a
        process group every rank saves ``<path>.rank<r>`` and rank 0 merges the shards in rank order."""
        if self._fill:
            self._drain(self._fill)
        self._side.synchronize()
        x = torch.cat(self._host, dim=0).contiguous() if self._host else torch.empty(0, self.dim, dtype=self.dtype)
        y = torch.cat(self._ys, dim=0).long().contiguous() if self._ys else torch.empty(0, dtype=torch.int64)
        if path is not None:
            path = Path(path)
            path.parent.mkdir(parents=True, exist_ok=True)
            rank, world = _world(self.group)
            if world == 1:
                torch.save({"x": x, "y": y}, path)
            else:
                torch.save({"x": x, "y": y}, f"{path}.rank{rank}")
                dist.barrier(group=self.group)
                if rank == 0:
                    merge_feature_shards(path, world)
                dist.barrier(group=self.group)
        return x, y


def merge_feature_shards(path, world: int, cleanup: bool = True) -> None:
    parts = [torch.load(f"{path}.rank{r}", map_location="cpu") for r in range(world)]
    blob = {"x": torch.cat([p["x"] for p in parts], dim=0).contiguous(),
            "y": torch.cat([p["y"] for p in parts], dim=0).long().contiguous()}
    torch.save(blob, path)
    if cleanup:
        for r in range(world):
            os.remove(f"{path}.rank{r}")


def load_feature_cache(path) -> Tuple[torch.Tensor, torch.Tensor]:
    """(features [N, D], labels [N] int64) from a cache file, accepting the key spellings the reference accepts
    (singletask_trainer.py:233-242: x | features | feats, y | labels)."""
    blob = torch.load(path, map_location="cpu")
    x_key = next((k for k in ("x", "features", "feats") if k in blob), None)
    y_key = next((k for k in ("y", "labels") if k in blob), None)
    if x_key is None or y_key is None:
        raise KeyError(f"Cache feature non riconosciuta: chiavi presenti {list(blob.keys())}")
    return blob[x_key].contiguous(), blob[y_key].long().contiguous()


def ensure_features(model, img_dataset, split: str, features_dir, task: str, batch_size: int = 64, num_workers: int = 0,
                    group=None):
    """Drop-in for ``SingleTaskTrainer._ensure_features`` (singletask_trainer.py:219-285): returns
    ``TensorDataset(features, y)`` for the split, from the cache file when present, else by running
    ``model.extract_features`` over this rank's contiguous share of the dataset and writing the (merged) cache."""
    from torch.utils.data import DataLoader, Subset, TensorDataset

    from .parallel import shard_range

    fpath = Path(features_dir) / f"{split}_features.pt"
    rank, world = _world(group)
    have = fpath.exists()
    if world > 1:
        # ONE decision for the whole group (rank 0's view of the file system): ranks that disagree — a stale or partially
        # visible cache file — would otherwise split between loading and the writer's barriers and hang each other
        flag = [have]
        dist.broadcast_object_list(flag, src=0, group=group)
        have = bool(flag[0])
    if not have:
        lo, hi = shard_range(len(img_dataset), rank, world)
        was_training = model.training
        model.eval()
        writer = None
        if hi > lo:
            loader = DataLoader(Subset(img_dataset, range(lo, hi)), batch_size=batch_size, shuffle=False, num_workers=num_workers,
                                drop_last=False, collate_fn=lambda b: ([s[0] for s in b], [s[1] for s in b]))
            with torch.no_grad():
                for images_list, targets_list in loader:
                    feats = model.extract_features(images_list)
                    if writer is None:
                        writer = FeatureCacheWriter(feats.shape[1], feats.device, group=group)
                    writer.add(feats.float(), targets_to_matrix(targets_list, [task])[:, 0])
        if writer is None:
            if world == 1:
                raise ValueError("empty dataset: nothing to cache")
            # fewer samples than ranks: this rank contributes an empty shard (and still meets the writer's barriers)
            bb = model.backbone
            dev = bb._tower_device() if hasattr(bb, "_tower_device") else next(model.parameters()).device
            writer = FeatureCacheWriter(int(bb.output_dim), dev, group=group)
        writer.finalize(fpath)
        if was_training:
            model.train()
    feats, ys = load_feature_cache(fpath)
    return TensorDataset(feats, ys)


@torch.no_grad()
def run_tester(model, loader, tasks: Sequence[str], classes: Union[int, Mapping[str, int]], out_dir=None,
               dataset_name: str = "MiviaPar", group=None) -> dict:
    """The body of ``BaseTester.run_one`` (base_tester.py:40-61) on the sink: ``model.predict`` per batch, no host sync
    until the loader is exhausted, then preds.json / gts.json / metrics.json."""
    sink = None
    model.eval()
    for images_list, targets_list in loader:
        pred = model.predict(images_list)
        if sink is None:
            dev = next(iter(pred.values())).device if isinstance(pred, Mapping) else pred.device
            sink = PredictionSink(tasks, classes, dev, group=group)
        sink.add(pred, targets_list)
    if sink is None:
        raise ValueError("empty loader")
    return sink.save(out_dir, dataset_name) if out_dir is not None else sink.metrics(dataset_name)
