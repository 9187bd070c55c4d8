#!/bin/bash
# ncu --set full of the tile-route walk kernels of one 1080p step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_compositor.py -x -q -m gpu -k "long_lists or pair_list" > gpurun_out/comp2.log 2>&1; echo "binning tests rc=$?"; tail -3 gpurun_out/comp2.log
python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_c3_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_view_backward|k_view_render|k_view_reduce|k_view_pairs' --launch-skip 8 -c 5 -f -o gpurun_out/r02_tile_walks python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_tile_walks.log 2>&1
ls -la gpurun_out/*.ncu-rep
