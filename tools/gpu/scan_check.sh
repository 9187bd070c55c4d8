#!/bin/bash
# scan-op parity + timings (C3 bench line without the side legs, C2 leg through the full bench)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/bench_scan.json 2> gpurun_out/bench_scan.err; echo "rc=$?"
python bench.py --workload c4 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/bench_scan_c4.json 2>> gpurun_out/bench_scan.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_scan.json", "gpurun_out/bench_scan_c4.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), "fwd", round(d["fwd_ms"], 4), round(d["roofline_fwd"]["frac"], 3), "bwd", round(d["bwd_ms"], 4), round(d["roofline"]["frac"], 3))
PY
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
print(bench.c2_leg(torch.device("cuda", 0))["scan_on_view0"])
PY
