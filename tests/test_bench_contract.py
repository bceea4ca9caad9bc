"""The bench lines committed under profiles/ (printed by bench.py on a B200) against the driver's JSON contract: required keys
and types, internal consistency (value = images / time, roofline fraction = achieved / peak, e2e copies declared), the
reference arm's line. CPU-only: guards the contract when bench.py is edited without a GPU at hand."""
import glob
import json
import math
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


TOWER_LINES = ["r2_y_bench_siglip_f16.json", "r2_y_bench_siglip_bf16.json", "r2_y_bench_clip.json", "r2_y_bench_eva.json",
               "r2_y_bench_siglip_n2.json"]


@pytest.mark.parametrize("name", TOWER_LINES)
def test_tower_bench_line_contract(name):
    d = _line(name)
    for k, t in [("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                 ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                 ("config", dict), ("e2e", dict), ("gpu_launches", int), ("clocks", dict), ("roofline", dict)]:
        assert isinstance(d[k], t), (k, type(d[k]))
    assert d["vs_baseline"] is None and BASELINE["published"] == {}  # no published number for this metric
    assert d["metric"] in BASELINE["metric"] and d["unit"] == "images/s" and d["higher_is_better"] and d["scaling"] == "weak"
    assert d["warmup"] >= 3 and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    # whole-job throughput = global batch / step time
    assert math.isclose(d["value"], d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3), rel_tol=2e-3)
    assert d["config"]["parallelism"] == f"dp{d['n_gpus']}"
    # end to end: copies declared and counted from the tensors (uint8 NHWC batch in, labels out)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0 and d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"]
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and math.isclose(r["frac"], r["achieved"] / r["peak"], rel_tol=1e-3)
    assert 0.5 < r["frac"] < 1.0 and 0.4 < r["step_frac_burst"] < r["frac"]
    assert d["config"]["tolerance_ok"] == (d["dtype"] == "fp16")  # the bf16 mode is fenced, not claimed


def test_headline_line_carries_cpu_baseline_and_traffic():
    d = _line("r2_y_bench_siglip_f16.json")
    assert "SigLIP-So400m" in d["config"]["workload"] and d["config"]["global_batch"] == 256 and d["n_gpus"] == 1
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["unit"] == d["unit"] and b["sample"] and b["cpu"]
    r = d["roofline"]
    assert r["traffic"] and 0.8 < r["traffic"] / r["algorithmic_bytes"] < 1.1  # no wasted re-reads in the dominant kernel
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
        rt = json.load(f)
    assert os.path.exists(os.path.join(ROOT, rt["source"].split(" ")[0]))  # the capture it cites is committed


def test_reference_arm_line():
    d = _line("r2_z_bench_reference.json")
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_probe_train_lines():
    for name in ("r2_y_pt_n2_p2p.json", "r2_z_pt_n1.json"):
        d = _line(name)
        assert d["unit"] == "steps/s" and d["higher_is_better"] and d["scaling"] == "weak"
        assert math.isclose(d["value"], 1.0 / (d["ms_per_step"] * 1e-3), rel_tol=5e-3)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["roofline"]["bound"] == "hbm"


def test_every_committed_json_profile_parses():
    for path in glob.glob(os.path.join(ROOT, "profiles", "*.json")):
        with open(path) as f:
            txt = f.read().strip()
        json.loads(txt.splitlines()[-1] if path.endswith(("_n2.json", "_n8.json")) else txt if txt.startswith("{\n") else txt.splitlines()[-1])
