#!/bin/bash
# compositor + batch + full-size tests, per-kernel breakdown of both scenes, a short bench (no CPU / reference legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_compositor.py tests/test_views_native.py tests/test_gpu_fullsize.py tests/test_abi.py -x -q -m gpu > gpurun_out/comp.log 2>&1; echo "compositor tests rc=$?"; tail -3 gpurun_out/comp.log
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --steps 6 2>&1 | grep "k_view\|splat step" | cut -c1-70,150-215 > gpurun_out/tile_breakdown.log
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 0 --steps 6 2>&1 | grep "k_view\|splat step" | cut -c1-70,150-215 >> gpurun_out/tile_breakdown.log
cat gpurun_out/tile_breakdown.log
timeout 900 python bench.py --no-reference-legs --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; grep "multi-view" gpurun_out/bench_q.err; tail -2 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
s=d["splat_step"]; print({k:s[k] for k in ("fwd_ms","bwd_ms","ms","native_call_ms","native_call_launches")}, s["e2e_host_tables"], s["multi_view"]["step_ms"], s.get("multi_view_python_loop"))
print([ (v["splat_ms"]) for v in s["c2_bundled"]["views"]])
PY
