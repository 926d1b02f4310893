"""The benchmark line's contract (task brief, "Measurement"): keys and types of the committed lines that
`bench.py` printed on a B200 -- a guard against format drift, not a measurement."""
import json
import os

import pytest

from conftest import ROOT

P = os.path.join(ROOT, "profiles")


def _load(name):
    # (the 4-GPU capture of round 1 still has NCCL's version banner in front of the line)
    lines = [ln for ln in open(os.path.join(P, name)).read().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_headline_line_has_the_contract_keys():
    d = _load("r01_bench_n1_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "hard_sphere_trial_moves_per_sec" and d["unit"] == "moves/s" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # whole-job throughput = units / time
    n = d["config"]["N"] * d["config"]["sweeps_per_step"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9


def test_reference_arm_line():
    d = _load("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "reference"
    mine = _load("r01_bench_n1_final.json")
    assert d["metric"] == mine["metric"] and d["unit"] == mine["unit"] and d["config"]["workload"] == mine["config"]["workload"]


@pytest.mark.parametrize("name", ["r01_bench_n4_fused.json", "r01_bench_widom_n1.json"])
def test_other_lines_parse(name):
    d = _load(name)
    assert d["value"] > 0 and d["unit"] and d["roofline"]["peak"] > 0
