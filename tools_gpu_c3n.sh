#!/bin/bash
mkdir -p gpurun_out
python tools/splat_time.py --steps 1 && ncu --set full --clock-control none --import-source on -k regex:'k_splat|k_place' --launch-skip 26 -c 13 -f -o gpurun_out/c3_full1 python tools/splat_time.py --steps 1 > gpurun_out/ncu_c3_full1.log 2>&1
ls -la gpurun_out/c3_full1.ncu-rep
