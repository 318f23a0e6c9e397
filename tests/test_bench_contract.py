"""The JSON lines bench.py prints are a contract with the driver: the committed lines of the final evidence pass
(profiles/r4z_*) carry every key the contract names, with sane types and mutually consistent numbers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.loads([l for l in fh if l.startswith("{")][0])


def test_product_line_carries_the_contract():
    d = _line("r4z_bench_pong_tc3.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"] and d["unit"] and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    rows, iters = d["config"]["rows_per_gpu"], d["config"]["iters_per_step"]
    assert abs(d["value"] - rows * iters / (d["ms_per_step"] / 1e3)) <= 1e-3 * d["value"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                         # measured on its own path, not a copy of the resident number
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and 0 < r["frac"] < 1.2 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("reference", "port") and c["sample"]
    assert d["gae"]["roofline"]["frac"] < 1.2
    for k in ("navlaser", "navimg"):
        assert d["other_workloads"][k]["value"] > 0


def test_reference_line_carries_the_contract():
    d, ours = _line("r4z_bench_reference_cpu.json"), _line("r4z_bench_pong_tc3.json")
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
    assert d["higher_is_better"] is True and d["config"]["workload"] == ours["config"]["workload"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_reference_variant_has_no_reference_arm():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "navlaser3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "unavailable" in d
