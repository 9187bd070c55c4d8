#!/bin/bash
# ncu --set full of the tile-route kernels of one 1080p step (all kernels of the view, one launch each)
mkdir -p gpurun_out
python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_c3_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_view_' --launch-skip 28 -c 14 -f -o gpurun_out/r02_tile_c3 python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
