#!/bin/bash
# record of the final tree of round 2 (after the radix binning / dependent launches): smoke, all GPU tests, the default
# bench line and its reference arm, launch lists (bench command; one tile-route view with DRAM bytes and
# warp-instructions), the tile-route view under ncu --set full, per-kernel breakdown of both scenes
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --no-reference-legs --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "bench list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_splat_tiles.csv python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_splat.log 2>&1; echo "splat list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_view_' --launch-skip 32 -c 16 -f -o gpurun_out/r02_tile_view python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_view.log 2>&1; echo "ncu tile rc=$?"
for c2 in -1 0; do
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 $c2 --steps 6 2>&1 | grep "k_view\|splat step\|elements" | cut -c1-70,150-215
done > gpurun_out/tile_breakdown.log 2>&1
