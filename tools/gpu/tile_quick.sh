#!/bin/bash
# quick check of the tile route: its bit-exact / oracle tests, then per-kernel device times of both scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py tests/test_views_native.py -x -q -m gpu -k "radix or pair_list or binning or fixture or ragged or edge or batch or lanes or split" > gpurun_out/tile_quick.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tile_quick.log
for c2 in -1 0; do
SPLAT_PROFILE=1 timeout 300 python tools/splat_time.py --route tiles --c2 $c2 --steps 6 2>&1 | grep "k_view\|elements" | cut -c1-70,150-215
done > gpurun_out/tile_breakdown.log 2>&1
grep -v "k_view_\(backward\|reduce\|cnt\|combine\|render\)" gpurun_out/tile_breakdown.log
