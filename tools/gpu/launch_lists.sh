#!/usr/bin/env bash
# ncu launch lists (device time per launch, cold-cache and serialised) of the bench command and of one splat step
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/bench_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
echo "bench list rc=$?"
python tools/splat_time.py --route tiles --steps 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_splat_tiles.csv python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_splat.log 2>&1
echo "splat list rc=$?"
wc -l gpurun_out/launches_bench.csv gpurun_out/launches_splat_tiles.csv
