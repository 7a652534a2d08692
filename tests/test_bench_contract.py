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
    d = _line("r2_bench_n1.json")
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
    # round 2: the reference's own CUDA timed in the same run, on the same scene
    rc = d["reference_cuda"]
    assert rc["speedup_of_this_library"] > 1 and rc["image_max_abs_err"] < 1e-4
    assert max(rc["grad_max_rel_err"].values()) < 1e-3
    assert "full config B" in c["sample"]  # the CPU arm runs the benchmark's own configuration
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_reference_arm_line():
    r = _line("r2_bench_n1_reference.json")
    assert r["impl"] == "reference" and r["unit"] == "Mpix/s" and r["metric"] == _line("r2_bench_n1.json")["metric"]
    assert r["config"]["workload"] == _line("r2_bench_n1.json")["config"]["workload"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["e2e"]["value"] == r["value"] and r["cpu_baseline"]["value"] == r["value"]


def test_committed_multi_gpu_lines_carry_parity_and_exchange():
    """N > 1 lines: gradients of the rank-sharded step equal the single-process batch (measured in the same run),
    the timed exchange is named, and the other exchanges are timed beside it."""
    base = _line("r2_bench_n1.json")["value"]
    for n in (2, 4, 8):
        d = _line(f"r2_bench_n{n}.json")
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert d["dp_parity_max_rel"] is not None and d["dp_parity_max_rel"] < 1e-3
        ex = d["exchange"]
        assert ex["timed"].startswith("peer") and ex["peer_unavailable"] is None
        assert "nccl_plain" in ex["ms_per_step_other"]
        assert d["value"] > 0.75 * n * base  # at least 0.75 of linear at every N
        assert d["config_C"]["ms_per_step"] > 0 and d["config_E"]["ms_per_step"] > 0
