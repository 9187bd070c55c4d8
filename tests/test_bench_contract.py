"""CPU: the bench lines committed under profiles/ (written by bench.py on a B200) carry every key of the driver's
contract, and their derived numbers are consistent with each other (roofline = algorithmic bytes / kernel time / peak,
value = elements / time, the reference arm mirrors the config)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_default_bench_line_has_the_contract_keys_and_consistent_numbers():
    d = _line("r02_bench_c3.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Gelem/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    n = d["config"]["elements_per_gpu"] * d["n_gpus"]
    assert d["value"] == pytest.approx(n / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    # achieved = SURVEY.md §8d's algorithmic bytes of the backward (16 N + 4 K) over its measured time
    alg = 16 * d["config"]["elements_per_gpu"] + 4 * d["config"]["segments_per_gpu"]
    assert r["algorithmic_bytes"] == alg
    assert r["achieved"] == pytest.approx(alg / (d["bwd_ms"] * 1e-3) / 1e9, rel=1e-6)
    assert 0.0 < r["frac"] < 1.0 and (r["traffic"] is None or r["traffic"] > 0.5 * alg)
    assert d["bwd_ms"] < d["ms_per_step"] and d["fwd_ms"] + d["bwd_ms"] == pytest.approx(d["ms_per_step"], rel=0.05)
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_mirrors_the_config():
    d, r = _line("r02_bench_c3.json"), _line("r02_bench_c3_reference_arm.json")
    assert r["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better", "config", "dtype"):
        assert r[k] == d[k], k
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["cores"] >= 1


@pytest.mark.parametrize("name,n", [("r02_bench_c3_2gpu.json", 2), ("r02_bench_c3_8gpu.json", 8)])
def test_multi_gpu_lines_are_whole_job_aggregates(name, n):
    d, one = _line(name), _line("r02_bench_c3.json")
    assert d["n_gpus"] == n and d["config"]["elements_per_gpu"] == one["config"]["elements_per_gpu"]
    # (every rank scans its own view: the ranks' element counts differ by a few 1e-4 from rank 0's)
    assert d["value"] == pytest.approx(n * d["config"]["elements_per_gpu"] / (d["ms_per_step"] * 1e-3) / 1e9, rel=2e-3)
    assert 0.9 * n * one["value"] < d["value"] < 1.1 * n * one["value"]       # weak scaling, no data-path collective


def test_reference_arm_runs_without_a_gpu_and_prints_one_contract_line():
    """`bench.py --impl reference` is CPU code (the pure-PyTorch path BASELINE.json names, oracle/torch_cpu_path.py):
    on the 1 Mi-element workload it finishes in seconds here; non-zero ranks print nothing and exit 0."""
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
           "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT,
                         env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["unit"] == "Gelem/s" and r["value"] > 0 and r["gpu_launches"] == 0
    assert r["config"]["elements_per_gpu"] == 1 << 20 and r["config"]["segments_per_gpu"] == 1 << 16
    assert r["e2e"] == {"value": r["value"], "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT,
                           env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert other.returncode == 0 and other.stdout.strip() == ""
