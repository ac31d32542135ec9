"""bench.py's output contract (no GPU): the reference arm really runs here on the CPU (one bounded step) and prints ONE JSON
line with the agreed keys; the committed bench lines under profiles/ carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "LR event-frames/sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_committed_bench_lines_have_every_key():
    for name in ("r1_bench_cfg2_n1.json", "r1_bench_cfg2_n2.json", "r1_bench_cfg2_n4.json", "r1_bench_cfg3_n1.json", "r1_bench_cfg4_n1.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert (BASE_KEYS - {"cpu_baseline"}) <= set(d), name
        assert {"clocks", "gpu_launches", "roofline"} <= set(d), name
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]), name
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"], name
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]), name
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9, name
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not set(d["clocks"]["reasons"]) & {
            "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}, name
        assert d["gpu_launches"] > 0 and d["config"]["workload"], name
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_cfg2_n1.json")))
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
