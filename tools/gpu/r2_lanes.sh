#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 2" "4 2" "2 2" "3 1"; do
set -- $cfg
GCP_WALK_PER_SM=$2 BENCH_SKIP_PY_LOOP=1 BENCH_MV_LANES=$1 timeout 900 python bench.py --no-reference-legs --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_mv_x.json 2> gpurun_out/bench_mv_x.err; echo "lanes=$1 walk CTAs/SM=$2 rc=$?"; grep "multi-view" gpurun_out/bench_mv_x.err | tail -1
done
