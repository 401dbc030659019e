"""CPU: the bench lines committed under profiles/r2/ (written by bench.py on a B200) carry every key of the driver's contract, and
both arms describe the same workload; bench.py's command line has the contract's flags and defaults."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R2 = os.path.join(ROOT, "profiles", "r2")


def _line(name):
    return json.loads(open(os.path.join(R2, name)).read().strip().splitlines()[-1])


def test_our_arm_line_has_the_contract_keys():
    d = _line("r2_bench_c384_A.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["unit"] == "cell-updates/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 1e9 and 0 < d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] == "port" and c["cores"] >= 1
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = cells * n_split / time
    cells = 6 * 384 * 384 * 79 * d["config"]["n_split"]
    assert abs(d["value"] - cells / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9


def test_reference_arm_line_matches_the_workload_of_our_arm():
    a, r = _line("r2_bench_c384_A.json"), _line("r2_bench_c384_reference.json")
    assert r["impl"] == "reference" and r["metric"] == a["metric"] and r["unit"] == a["unit"] and r["higher_is_better"] == a["higher_is_better"]
    assert r["config"]["workload"] == a["config"]["workload"]
    assert r["e2e"]["value"] == r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["kind"] == "port"
    assert r["steps"] >= 1 and a["value"] / r["value"] > 20.0          # the north star's >= 20x, against the honest arm


def test_bench_command_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out, flag
