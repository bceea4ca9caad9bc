"""Tester sink and feature cache: metrics / file formats against the reference Evaluator's own
outputs (tests/golden/evaluator.json, made by tests/golden/make_golden_evaluator.py), the GPU confusion counts against
sklearn, and the sharded cache writer."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from make_golden_evaluator import CASES, seeded_labels  # noqa: E402

from vlmfw_b200 import sinks  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "evaluator.json")))


def _counts(pred, gt, cmax):
    counts = np.zeros((pred.shape[1], cmax, cmax), np.int64)
    for j in range(pred.shape[1]):
        ok = gt[:, j] >= 0
        np.add.at(counts[j], (gt[ok, j], pred[ok, j]), 1)
    return counts


# ------------------------------------------------------------------------------------------------ CPU: host logic
@pytest.mark.parametrize("name", list(CASES))
def test_metrics_from_counts_equal_reference_evaluator(name):
    ds, tasks, rows, seed, missing = CASES[name]
    pred, gt = seeded_labels(tasks, rows, seed, missing)
    m = sinks.metrics_from_counts(list(tasks), _counts(pred, gt, max(tasks.values())), ds)
    assert json.dumps(m, indent=4) == GOLD[name]["metrics.json"]


def test_targets_to_matrix_and_cache_keys(tmp_path):
    t = sinks.targets_to_matrix([{"gender": 1, "age": 3}, {"gender": None}, {"age": "7"}], ["gender", "age"])
    assert t.tolist() == [[1, 3], [-1, -1], [-1, 7]]
    x, y = torch.randn(5, 3), torch.arange(5, dtype=torch.int32)
    for kx, ky in (("x", "y"), ("features", "labels"), ("feats", "y")):
        torch.save({kx: x, ky: y}, tmp_path / "c.pt")
        gx, gy = sinks.load_feature_cache(tmp_path / "c.pt")
        assert torch.equal(gx, x) and gy.dtype == torch.int64 and torch.equal(gy, y.long())
    torch.save({"a": x}, tmp_path / "bad.pt")
    with pytest.raises(KeyError):
        sinks.load_feature_cache(tmp_path / "bad.pt")


def test_merge_feature_shards_restores_rank_order(tmp_path):
    from vlmfw_b200.parallel import shard_range

    x, y = torch.randn(11, 4), torch.arange(11)
    for r in range(3):
        lo, hi = shard_range(11, r, 3)
        torch.save({"x": x[lo:hi], "y": y[lo:hi]}, f"{tmp_path / 'train_features.pt'}.rank{r}")
    sinks.merge_feature_shards(tmp_path / "train_features.pt", 3)
    gx, gy = sinks.load_feature_cache(tmp_path / "train_features.pt")
    assert torch.equal(gx, x) and torch.equal(gy, y) and sorted(os.listdir(tmp_path)) == ["train_features.pt"]


def test_sinks_need_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sinks.PredictionSink(["a"], 2, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sinks.FeatureCacheWriter(8, "cpu")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_prediction_sink_files_equal_reference_evaluator(cuda_device, tmp_path, name):
    ds, tasks, rows, seed, missing = CASES[name]
    pred, gt = seeded_labels(tasks, rows, seed, missing)
    names = list(tasks)
    sink = sinks.PredictionSink(names, tasks, cuda_device, capacity=16)  # forces the buffers to grow
    for lo in range(0, rows, 37):  # ragged last batch; alternate the accepted input forms
        p = torch.from_numpy(pred[lo:lo + 37]).cuda()
        g = gt[lo:lo + 37]
        if (lo // 37) % 2:
            sink.add({t: p[:, j] for j, t in enumerate(names)}, [{t: int(v) for t, v in zip(names, row)} for row in g])
        else:
            sink.add(p.to(torch.int32) if p.shape[1] > 1 else p[:, 0], torch.from_numpy(g))
    m = sink.save(tmp_path, ds)
    assert open(tmp_path / "metrics.json").read() == GOLD[name]["metrics.json"]
    assert open(tmp_path / "preds.json").read()[:400] == GOLD[name]["preds.json"]
    assert open(tmp_path / "gts.json").read()[:400] == GOLD[name]["gts.json"]
    from sklearn.metrics import confusion_matrix

    for j, t in enumerate(names):
        ok = gt[:, j] >= 0
        want = confusion_matrix(gt[ok, j], pred[ok, j], labels=list(range(tasks[t])))
        np.testing.assert_array_equal(sink.confusion_matrix(t), want)
    assert sink.finalize()["dropped"] == 0 and m == json.loads(GOLD[name]["metrics.json"])


@pytest.mark.gpu
def test_prediction_sink_on_probe_predictions(cuda_device):
    """End to end: MultiTaskProbe.predict on the GPU -> sink -> the same preds list parsed_attributes builds per batch."""
    from vlmfw_b200 import probes, specs, weights
    from vlmfw_b200.vision_backbone import backbone_from_spec

    spec = specs.CLIP_TINY
    bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, 3), device="cuda")
    tasks = {"gender": 2, "upper": 12, "bag": 2}
    probe = probes.MultiTaskProbe(bb, tasks).cuda().eval()
    imgs = weights.seeded_images_u8(10, spec.image_size, 5)
    targets = [{"gender": i % 2, "upper": i % 12, "bag": None if i % 3 == 0 else 1} for i in range(10)]
    loader = [(imgs[:6], targets[:6]), (imgs[6:], targets[6:])]
    sink = sinks.PredictionSink(list(tasks), tasks, cuda_device)
    want = []
    for im, tg in loader:
        pred = probe.predict(im)
        sink.add(pred, tg)
        want += probes.parsed_attributes(pred)
    preds, gts = sink.preds_and_gts()
    assert preds == want and [g["bag"] for g in gts] == [-1 if i % 3 == 0 else 1 for i in range(10)]
    assert sinks.run_tester(probe, loader, list(tasks), tasks) == sink.metrics()


@pytest.mark.gpu
def test_feature_cache_writer_streams_chunks(cuda_device, tmp_path):
    w = sinks.FeatureCacheWriter(24, cuda_device, chunk_rows=16)
    g = torch.Generator().manual_seed(0)
    xs, ys = [], []
    for b in (5, 16, 1, 40, 3):  # crosses chunk boundaries, several drains in flight
        x = torch.randn(b, 24, generator=g)
        y = torch.randint(-1, 9, (b,), generator=g)
        w.add(x.cuda(), y if b % 2 else y.tolist())
        xs.append(x)
        ys.append(y)
    x, y = w.finalize(tmp_path / "feat" / "val_features.pt")
    assert torch.equal(x, torch.cat(xs)) and torch.equal(y, torch.cat(ys))
    blob = torch.load(tmp_path / "feat" / "val_features.pt")
    assert set(blob) == {"x", "y"} and torch.equal(blob["x"], x) and blob["y"].dtype == torch.int64


@pytest.mark.gpu
def test_ensure_features_matches_per_batch_extraction(cuda_device, tmp_path):
    """Drop-in for SingleTaskTrainer._ensure_features: same file, same tensors as the reference's per-batch .cpu() loop."""
    from vlmfw_b200 import probes, specs, weights
    from vlmfw_b200.vision_backbone import backbone_from_spec

    spec = specs.SIGLIP_TINY
    bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, 1), device="cuda")
    probe = probes.LinearProbe(bb, 9).cuda()
    imgs = weights.seeded_images_u8(7, spec.image_size, 2)
    ds = [(imgs[i], {"age": i % 9 if i != 4 else None}) for i in range(7)]
    td = sinks.ensure_features(probe, ds, "train", tmp_path, "age", batch_size=3)
    probe.eval()
    want = torch.cat([probe.extract_features(imgs[i:i + 3]).cpu() for i in range(0, 7, 3)])
    assert torch.equal(td.tensors[0], want) and td.tensors[1].tolist() == [0, 1, 2, 3, -1, 5, 6]
    again = sinks.ensure_features(probe, ds, "train", tmp_path, "age")  # served from the cache file
    assert torch.equal(again.tensors[0], want)
