#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "launchlist rc=$?"
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 900 python bench.py --workload c4 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 900 python bench.py --workload c1 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"
python - <<PY
import json
for w in ("", "_c4", "_c1"):
    d=json.loads(open(f"gpurun_out/bench{w}.json").read().strip().splitlines()[-1])
    print(w or "c3", "value", round(d["value"],1), "fwd", round(d["fwd_ms"],4), "bwd", round(d["bwd_ms"],4), "fwd frac", round(d["roofline_fwd"]["frac"],3), "bwd frac", round(d["roofline"]["frac"],3), "both", round(d["roofline_fwd_bwd"]["frac"],3), "e2e", round(d["e2e"]["value"],2))
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1]); print(d["splat_step"]["ms"], d["splat_step"]["multi_view"]["step_ms"])
PY
