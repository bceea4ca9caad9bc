"""Generate tests/golden/evaluator.json by running THE REFERENCE's Evaluator (datasets_vlm/evaluate_dataset.py, imported
from /root/reference, unmodified) on seeded predictions / ground truths, the way BaseTester.run_one feeds it
(probing/test/base_tester.py:53-60). matplotlib is not installed: a no-op stub stands in for the plotting calls.

    PYTHONPATH=/root/reference python tests/golden/make_golden_evaluator.py
"""
import json
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = {
    # name: (dataset_name, {task: classes}, rows, seed, missing fraction)
    "mivia": ("MiviaPar", {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}, 257, 1, 0.15),
    "face_age": ("UTKFace", {"age": 9}, 100, 2, 0.0),
    "face_multi": ("FairFace", {"gender": 2, "ethnicity": 4, "emotion": 7, "age": 9}, 301, 3, 0.3),
    "single_class_seen": ("MiviaPar", {"bag": 2}, 40, 4, 0.5),
}


def seeded_labels(tasks, rows, seed, missing):
    """pred [rows, T] in [0, C_t), gt likewise with -1 where missing; predictions agree with truth ~60 % of the time."""
    rng = np.random.default_rng(seed)
    pred = np.zeros((rows, len(tasks)), np.int64)
    gt = np.zeros((rows, len(tasks)), np.int64)
    for j, c in enumerate(tasks.values()):
        gt[:, j] = rng.integers(0, c, rows)
        pred[:, j] = np.where(rng.random(rows) < 0.6, gt[:, j], rng.integers(0, c, rows))
        gt[rng.random(rows) < missing, j] = -1
    if len(tasks) == 1 and missing >= 0.5:
        pred[:] = 1  # a class that only ever appears as a prediction still shows up in "labels"
        gt[gt >= 0] = 0
    return pred, gt


def main():
    sys.path.insert(0, REF)
    plt = mock.MagicMock()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    from datasets_vlm.evaluate_dataset import Evaluator  # noqa: E402 (reference)
    from datasets_vlm.face_dataset import FaceDataset  # noqa: E402 (reference)

    out = {}
    for name, (ds, tasks, rows, seed, missing) in CASES.items():
        assert ds == "MiviaPar" or ds in FaceDataset.get_available_datasets(), (ds, FaceDataset.get_available_datasets())
        pred, gt = seeded_labels(tasks, rows, seed, missing)
        names = list(tasks)
        preds = [{t: int(v) for t, v in zip(names, row)} for row in pred]
        gts = [{t: int(v) for t, v in zip(names, row)} for row in gt]
        with tempfile.TemporaryDirectory() as tmp:
            Evaluator.evaluate(preds, gts, output_dir=tmp, dataset_name=ds, age_mode="classification")
            out[name] = {f: open(os.path.join(tmp, f)).read() for f in ("metrics.json", "preds.json", "gts.json")}
        out[name]["preds.json"] = out[name]["preds.json"][:400]  # the format, not the bulk
        out[name]["gts.json"] = out[name]["gts.json"][:400]
    with open(os.path.join(HERE, "evaluator.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote evaluator.json")


if __name__ == "__main__":
    main()
