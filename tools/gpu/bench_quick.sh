#!/bin/bash
# the bench without its CPU / reference legs: splat step, batch of views, bundled views
mkdir -p gpurun_out
timeout 900 python bench.py --no-reference-legs --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; grep "multi-view" gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
s=d["splat_step"]; print({k:round(s[k],4) for k in ("fwd_ms","bwd_ms","ms","native_call_ms","native_call_launches")}, round(s["e2e_host_tables"]["ms"],3), round(s["multi_view"]["step_ms"],3), s.get("multi_view_python_loop",{}).get("ms_per_view"))
print([round(v["splat_ms"],4) for v in s["c2_bundled"]["views"]], d["value"], d["fwd_ms"], d["bwd_ms"], d["e2e"])
PY
