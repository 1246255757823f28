"""CPU: the committed bench lines (profiles/r02n_bench.json / r02n_bench_ref.json, written by `python bench.py [--impl reference]`
on a B200 box) carry every key of the bench contract, with consistent values."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_follows_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02n_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert "workload" in d["config"] and "4096" in d["config"]["workload"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    # value = envs * steps / time
    assert abs(d["value"] - 4096 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert d["steps_timed"] % 32 == 0 and d["steps_timed"] >= d["steps"]            # whole horizons inside the timed window
    assert d["post_horizon_share"]["charged_ms"] == 0.0 and d["post_horizon_share"]["passes_in_window"] == d["steps_timed"] // 32
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


def test_committed_train_step_and_reference_arm_lines():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02n_bench.json")))
    t = d["train_step"]
    assert t["metric"] == "train_samples_per_sec" and t["minibatch"] == 16384 and t["parameters"] > 11_000_000
    assert abs(t["value"] - t["minibatch"] / (t["ms_per_minibatch"] * 1e-3)) / t["value"] < 1e-6
    assert 0 < t["roofline"]["frac"] <= 1.0 / 3.0 and t["forward_backward_ms"] < t["ms_per_minibatch"]
    p = d["kernels"]["post_step"]
    assert p["bytes_incl_sinks"] > 28264 and p["frac_incl_sinks"] > p["frac"]
    r = json.load(open(os.path.join(ROOT, "profiles", "r02n_bench_ref.json")))
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]
    assert r["cpu_baseline"]["kind"] == "port" and "all 4096 envs" in r["cpu_baseline"]["sample"] and r["cpu_baseline"]["cores"] >= 1
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["value"] == r["value"]
    assert d["e2e"]["value"] / r["value"] > 100          # end-to-end speed-up over the CPU arm on the same box
