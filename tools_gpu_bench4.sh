#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd", round(d["fwd_ms"],4), "bwd", round(d["bwd_ms"],4), "frac", round(d["roofline_fwd_bwd"]["frac"],3))
print("cpu", d["cpu_baseline"]); print("cpu_c", d["cpu_baseline_c_omp"]); print("e2e", d["e2e"]); print("splat ms", d["splat_step"]["ms"])
PY
