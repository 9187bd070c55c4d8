#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench N=1 rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "fwd", round(d["fwd_ms"],4), "bwd", round(d["bwd_ms"],4), "frac", round(d["roofline_fwd_bwd"]["frac"],3))
print("splat", json.dumps(d["splat_step"]))
print("cpu", d["cpu_baseline"]["value"], "e2e", d["e2e"]["value"])
PY
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "n_gpus", d["n_gpus"])
print("splat", json.dumps(d["splat_step"]))
PY
