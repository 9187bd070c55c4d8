#!/usr/bin/env bash
# multi-view leg, N ranks, for several tail sizes:  bash tools/gpu/multi_mv.sh N "4 3"
set -u
mkdir -p gpurun_out
N=${1:-8}
for tail in ${2:-2}; do
BENCH_SKIP_PY_LOOP=1 BENCH_MV_TAIL=$tail timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$tail bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-reference-legs > gpurun_out/bench_mv_n${N}_t$tail.json 2> gpurun_out/bench_mv_n${N}_t$tail.err; echo "N=$N tail=$tail rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_mv_n${N}_t$tail.json").read().strip().splitlines()[-1])
m=d["splat_step"]["multi_view"]; print(m["step_ms"], m["exposed_after_last_view_ms"], m["host_enqueue_ms"])
PY
done
