"""CPU: the JSON line bench.py prints (a committed sample from a real B200 run, profiles/r1_bench_final_session.json)
carries every key the measurement contract names, with consistent values; and the reference arm / CLI parse."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_final_session.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"] and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "65536" in d["config"]["workload"]
    n = d["config"]["n_envs_per_gpu_step"]
    assert abs(d["value"] - d["n_gpus"] * n / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert d["gpu_launches"] == d["steps"] and d["warmup"] >= 3
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert abs(r["achieved"] - n * 349 / (d["ms_per_step"] * 1e-3) / 1e9) / r["achieved"] < 1e-6     # algorithmic bytes / launch time
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == n * 16 and e["d2h_bytes_per_step"] == n * (72 + 4 + 1) and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU: the oracle port on the host cores (a bounded sample)."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the other ranks of a torchrun launch exit quietly
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""
