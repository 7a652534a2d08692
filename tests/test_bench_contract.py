"""The bench line contract (no GPU): the last committed capture must carry every key the driver and
the judge read, with consistent values."""
import json
import os

from conftest import ROOT


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        for ln in f:
            if ln.startswith("{"):
                return json.loads(ln)
    raise AssertionError(name)


def test_committed_bench_line_has_the_contract_keys():
    d = _line("r1_f_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = pixels of the step / time of the step
    assert abs(d["value"] - 1920 * 1080 * d["n_gpus"] / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "Mpix/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # copies inside the timed region can only cost
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["traffic"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == "Mpix/s" and c["sample"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_reference_arm_line():
    r = _line("r1_f_bench_reference.json")
    assert r["impl"] == "reference" and r["unit"] == "Mpix/s" and r["metric"] == _line("r1_f_bench.json")["metric"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["e2e"]["value"] == r["value"] and r["cpu_baseline"]["value"] == r["value"]
