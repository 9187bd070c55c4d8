#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for w in c3 c4 c1; do
  timeout 900 python bench.py --workload $w --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", d["config"]["workload"], "n=",d["config"]["elements_per_gpu"], "value", round(d["value"],1), "fwd_ms", round(d["fwd_ms"],4), "bwd_ms", round(d["bwd_ms"],4), "fwd frac", round(d["roofline_fwd"]["frac"],3), "bwd frac", round(d["roofline"]["frac"],3), "both", round(d["roofline_fwd_bwd"]["frac"],3), d["clocks"])
PY
  tail -2 gpurun_out/bench_$w.err
done
