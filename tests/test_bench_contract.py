"""The bench.py JSON contract, checked on the committed lines of the last GPU runs (profiles/r02_*): the keys the driver
and the judge read must be present and consistent.  (bench.py itself needs a GPU; this guards the schema.)"""
import glob
import json
import os

import pytest

from conftest import ROOT


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    assert os.path.exists(path), path
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_our_arm_line_has_the_contract_keys():
    j = _line("r02_bench_default_goe_n16384_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "check"):
        assert k in j, k
    assert j["unit"] == "s" and j["higher_is_better"] is False and j["scaling"] == "strong" and j["dtype"] == "f64"
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and j["n_gpus"] == 1 and j["warmup"] >= 3
    assert abs(j["ms_per_step"] - j["value"] * 1e3) < 1e-9
    # headline = BASELINE configs[2] at every N; the config dict holds nothing but the workload (both arms print the same)
    assert set(j["config"]) == {"workload", "n", "matrix", "ref_leaves"} and j["config"]["n"] == 16384 and j["config"]["matrix"] == "goe"
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["frac"] > 0.9
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
    # parity verdict of the headline and of every extra configuration
    assert j["check"]["parity"] is True and j["check"]["parity_all_configs"] is True
    assert j["check"]["merge_stats_identical"] is True and j["check"]["lambda_max_abs_diff_vs_reference"] <= j["check"]["lambda_tol"]
    for name in ("s1_4k", "wilk16k"):
        x = j["other_configs"][name]
        assert x["check"]["parity"] is True and "roofline" in x and x["value"] > 0
    assert j["eigenvalues_only"]["value"] < j["value"]


def test_reference_arm_line():
    j = _line("r02_bench_reference_arm_goe_n16384.json")
    assert j["impl"] == "reference" and j["unit"] == "s" and j["higher_is_better"] is False and j["extrapolated"] is True
    assert j["e2e"] == {"value": j["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["value"] == j["value"]
    ours = _line("r02_bench_default_goe_n16384_1gpu.json")
    assert j["config"] == ours["config"] and j["metric"] == ours["metric"]
    # the fully measured pair: eigenvalue phase of the reference vs our eigenvalue-only solve
    assert j["eigenvalues_only"]["value"] > 1.0 and ours["eigenvalues_only"]["value"] < 0.1


@pytest.mark.parametrize("pattern", ["r02_bench_*_g2*.json", "r02_bench_*_g4*.json", "r02_bench_*_g8*.json"])
def test_multi_gpu_lines(pattern):
    for f in glob.glob(os.path.join(ROOT, "profiles", pattern)):
        j = json.loads(open(f).read().strip().splitlines()[-1])
        assert j["n_gpus"] > 1 and j["scaling"] == "strong" and j["value"] > 0 and j["gpu_launches"] > 0
        assert j["check"]["parity"] is True and j["config"]["n"] == 16384
