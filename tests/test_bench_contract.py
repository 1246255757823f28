"""CPU: the committed bench line (profiles/r01g_bench.json, written by `python bench.py` on a B200) carries every key of the
bench contract, with consistent values."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_follows_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01g_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert "workload" in d["config"] and "4096" in d["config"]["workload"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    # value = envs * steps / time
    assert abs(d["value"] - 4096 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert set(e) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                                   # host buffers in and out every step cannot be faster
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0 / 3.0 + 1e-9     # bf16x3 ceiling
    c = d["cpu_baseline"]
    assert set(c) >= {"value", "unit", "cores", "kind", "sample"} and c["kind"] in ("port", "reference") and c["cores"] >= 1
    assert c["unit"] == d["unit"] and 0 < c["value"] < d["value"]
    k = d["clocks"]
    assert set(k) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not any(x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for x in k["reasons"])
