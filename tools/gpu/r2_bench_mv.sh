#!/bin/bash
mkdir -p gpurun_out
for lanes in 1 2; do
BENCH_MV_LANES=$lanes timeout 900 python bench.py --no-reference-legs --no-cpu-baseline > gpurun_out/bench_mv_l$lanes.json 2> gpurun_out/bench_mv_l$lanes.err; echo "bench lanes=$lanes rc=$?"; grep "multi-view" gpurun_out/bench_mv_l$lanes.err; tail -2 gpurun_out/bench_mv_l$lanes.err
done
python - <<PY
import json
for l in (1,2):
    d=json.loads(open(f"gpurun_out/bench_mv_l{l}.json").read().strip().splitlines()[-1])
    s=d["splat_step"]; print(l, {k:s[k] for k in ("fwd_ms","bwd_ms","ms")}, s["multi_view"], s.get("multi_view_python_loop"))
PY
