#!/bin/bash
mkdir -p gpurun_out
python tools/splat_time.py --c2 1 --steps 1 && ncu --set full --clock-control none --import-source on -k regex:'k_splat|k_place' --launch-skip 26 -c 13 -f -o gpurun_out/c2_full3 python tools/splat_time.py --c2 1 --steps 1 > gpurun_out/ncu_c2_full3.log 2>&1
ls -la gpurun_out/c2_full3.ncu-rep
