#!/bin/bash
# ncu --set full of the tile-route kernels of one 1080p step and one bundled-scene step
mkdir -p gpurun_out
python tools/splat_time.py --route tiles --steps 1 && ncu --set full --clock-control none --import-source on -k regex:'k_view_render|k_view_backward|k_view_reduce' --launch-skip 3 -c 3 -f -o gpurun_out/tile_c3 python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_view_render|k_view_backward|k_view_reduce' --launch-skip 3 -c 3 -f -o gpurun_out/tile_c2 python tools/splat_time.py --route tiles --c2 1 --steps 1 > gpurun_out/ncu_tile_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 4K / 6 M Gaussians through both routes (SURVEY.md §8d, C4 splat route)
for r in tiles lists; do python tools/splat_time.py --route $r --n 6000000 --width 3840 --height 2160 --steps 5 2>&1 | tail -2; done > gpurun_out/splat_4k.log 2>&1
cat gpurun_out/splat_4k.log
