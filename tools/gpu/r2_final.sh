#!/bin/bash
# round-2 evidence run on one B200: tests + bench (default, C1, C4, alignment sweep), launch lists, ncu --set full of
# the default scan kernels and of every kernel of one tile-route view, per-kernel breakdowns, error table, 4K views
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload c1 --no-splat --no-reference-legs > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"
timeout 900 python bench.py --workload c4 --no-splat --no-reference-legs --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --sweep --steps 10 > gpurun_out/sweep.json 2> gpurun_out/sweep.txt; echo "sweep rc=$?"
# launch lists (ncu, serialised and cold-cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-splat --no-reference-legs --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "bench list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_splat_tiles.csv python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_splat.log 2>&1; echo "splat list rc=$?"
# ncu --set full: default scan kernels (C3), every kernel of one tile-route view (1080p)
for op in fwd bwd; do
ncu --set full --clock-control none --import-source on -k regex:k_${op}_blk -s 3 -c 1 -f -o gpurun_out/r02_scan_$op python tools/prof_one.py --op $op --variant -1 --halo 1 > gpurun_out/ncu_$op.log 2>&1; echo "ncu $op rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:'k_view_' --launch-skip 30 -c 15 -f -o gpurun_out/r02_tile_view python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_view.log 2>&1; echo "ncu tile rc=$?"
# per-kernel device times (kineto) of one step, both scenes
for c2 in -1 0; do
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 $c2 --steps 6 2>&1 | grep "k_view\|splat step\|elements" | cut -c1-70,150-215
done > gpurun_out/tile_breakdown.log 2>&1
python tools/compositor_errors.py --big > gpurun_out/compositor_errors.txt 2>&1; echo "error table rc=$?"
for r in tiles lists; do python tools/splat_time.py --route $r --n 6000000 --width 3840 --height 2160 --steps 5 2>&1 | tail -2; done > gpurun_out/splat_4k.log 2>&1
ls -la gpurun_out | head -50
