"""The bench lines kept under profiles/ (written by bench.py on a B200) carry every key the driver's contract names:
a format check of the committed records, no GPU and no bench run needed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.readline())


def test_repo_arm_line():
    d = _line("r02_bench_final_1gpu.json")
    assert BASE | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f32" and d["data"] == "synthetic"
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith(d["metric"].split(",")[0]) and d["unit"] == "agent-steps/s"      # agent-steps/sec ...
    assert "65,536" in d["metric"] and "TwoDBicycle" in d["metric"] and d["config"]["n_agents"] == 65536
    assert d["warmup"] >= 3 and "l2" in d["config"] and "workload" in d["config"]
    assert abs(d["value"] - d["config"]["n_agents"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    assert r["kernel_ms"] <= d["ms_per_step"] * 1.1          # the dominant kernel fits into the step
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert d["gpu_launches"] >= 3 * d["steps"]               # three launches per step (+ re-sorts)
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    repo = _line("r02_bench_final_1gpu.json")
    assert d["metric"] == repo["metric"] and d["unit"] == repo["unit"]


@pytest.mark.parametrize("name", ["r02_scaling_1_2_4_8.jsonl", "r02_scaling_final_build_1_2.jsonl"])
def test_scaling_lines(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [json.loads(l) for l in f if l.strip()]
    assert [l["n_gpus"] for l in lines] == sorted(l["n_gpus"] for l in lines)
    for l in lines:
        assert l["scaling"] == "strong" and l["config"]["n_agents"] == 65536
    assert all(b["value"] > a["value"] for a, b in zip(lines, lines[1:]))      # more GPUs, more agent-steps per second
