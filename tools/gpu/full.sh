#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd", round(d["fwd_ms"],4), "bwd", round(d["bwd_ms"],4), "frac", round(d["roofline_fwd_bwd"]["frac"],3), "bwd frac", round(d["roofline"]["frac"],3))
print("cpu", d["cpu_baseline"]["value"], "cpu_c", d["cpu_baseline_c_omp"]["value"], "e2e", d["e2e"]["value"], "splat ms", d["splat_step"]["ms"], "mv", d["splat_step"]["multi_view"]["step_ms"])
print("launches", d["gpu_launches"], d["clocks"])
PY
