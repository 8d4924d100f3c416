"""The bench.py JSON contract, checked on the committed lines of the last GPU runs (profiles/): the keys the driver
and the judge read must be present and consistent.  (bench.py itself needs a GPU; this guards the schema.)"""
import glob
import json
import os

import pytest

from conftest import ROOT


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return files[-1]


def test_our_arm_line_has_the_contract_keys():
    j = _line(_latest("r01_bench_default_s1_n4096_1gpu_v4.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in j, k
    assert j["unit"] == "s" and j["higher_is_better"] is False and j["scaling"] == "strong" and j["dtype"] == "f64"
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and j["n_gpus"] == 1 and j["warmup"] >= 3
    assert abs(j["ms_per_step"] - j["value"] * 1e3) < 1e-9
    assert "workload" in j["config"] and "model" not in j["config"] and "-s 1 -n 4096" in j["config"]["workload"]
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = j["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    e = j["e2e"]
    assert e["unit"] == "s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > j["value"]
    assert j["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in j["clocks"], k
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    j = _line(_latest("r01_bench_reference_arm_s1_n4096_v2.json"))
    assert j["impl"] == "reference" and j["unit"] == "s" and j["higher_is_better"] is False
    assert j["e2e"] == {"value": j["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["value"] == j["value"]
    ours = _line(_latest("r01_bench_default_s1_n4096_1gpu_v4.json"))
    assert j["config"]["workload"] == ours["config"]["workload"] and j["metric"] == ours["metric"]


@pytest.mark.parametrize("pattern", ["r01_scaleC_goe_*_g8.json", "r01_bench_default_s1_n4096_2gpu_with_config2.json"])
def test_multi_gpu_lines(pattern):
    for f in glob.glob(os.path.join(ROOT, "profiles", pattern)):
        j = _line(f)
        assert j["n_gpus"] > 1 and j["scaling"] == "strong" and j["value"] > 0 and j["gpu_launches"] > 0
        assert "sharding" in j["config"]
