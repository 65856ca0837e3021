"""The JSON line bench.py prints (driver contract): checked on the committed round-1 lines in profiles/ (GPU runs) and on
the argument surface of bench.py itself (no GPU needed)."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _check_line(d, reference=False):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "converged_game_solves_per_sec" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    cb = d["cpu_baseline"]
    assert set(cb) >= {"value", "unit", "cores", "kind", "sample"} and cb["kind"] in ("port", "reference")
    if reference:
        assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["e2e"]["h2d_bytes_per_step"] == 0
        return
    assert d["gpu_launches"] >= d["steps"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["warmup"] >= 3


def _last_line(name):
    return json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])


def test_committed_bench_lines_follow_the_contract():
    for name in ("r1_s15_bench.json", "r1_s13_bench_merge.json"):
        _check_line(json.loads((ROOT / "profiles" / name).read_text()))
    _check_line(json.loads((ROOT / "profiles" / "r1_s13_bench_ref.json").read_text()), reference=True)
    two = json.loads((ROOT / "profiles" / "r1_s12_bench_2gpu.json").read_text().strip().splitlines()[-1])
    assert two["n_gpus"] == 2 and two["config"]["instances_per_gpu"] == 2960


def test_round2_bench_lines_follow_the_contract():
    """The round-2 lines of every BASELINE workload: contract keys, the reference arm on the SAME config as the GPU arm
    (same_config was false), measured DRAM traffic for the two profiled workloads, single-instance latency."""
    head = _last_line("r2_close_bench.json")
    _check_line(head)
    ref = _last_line("r2_close_bench_ref.json")
    _check_line(ref, reference=True)
    assert ref["config"] == head["config"] and ref["metric"] == head["metric"] and ref["unit"] == head["unit"]
    assert head["roofline"]["traffic"] is not None and head["single_instance"]["gpu_ms_median"] > 0
    assert head["cpu_baseline"]["value"] > 5.0          # the BLAS-thread defect of round 1 gave 0.35
    for wl in ("merge", "curve", "agents3", "agents4"):
        d = _last_line(f"r2_close_bench_{wl}.json" if wl != "agents4" else "r2_final_bench_agents4.json")
        _check_line(d)
        assert d["config"]["workload"].startswith(wl[:5]) or wl.startswith("agents")
    two = _last_line("r2_close_bench_2gpu.json")
    assert two["n_gpus"] == 2 and 1.9 < two["value"] / head["value"] < 2.1


def test_bench_defaults_and_arguments():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
        assert flag in out
    src = (ROOT / "bench.py").read_text()
    assert 'add_argument("--gpus", type=int, default=1)' in src and 'add_argument("--warmup", type=int, default=3)' in src
