#!/bin/bash
# ncu --set full of the radix-binning kernels of one 1080p step
mkdir -p gpurun_out
python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_bin_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_view_pairs' --launch-skip 2 -c 1 -f -o gpurun_out/r02_bin python tools/splat_time.py --route tiles --steps 1 > gpurun_out/ncu_bin.log 2>&1
ls -la gpurun_out/*.ncu-rep
