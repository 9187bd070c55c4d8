#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamer" > gpurun_out/pytest_e2e.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_e2e.log
timeout 1200 python bench.py --no-splat --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_e2e.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e2e.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", d["e2e"])
PY
