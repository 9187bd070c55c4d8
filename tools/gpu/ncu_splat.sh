#!/bin/bash
mkdir -p gpurun_out
python tools/splat_time.py --c2 1 --steps 1 && ncu --set full --clock-control none --import-source on -k regex:'k_splat_bwd_elem_cells|k_place_fill_long' --launch-skip 4 -c 2 -f -o gpurun_out/c2_full4 python tools/splat_time.py --c2 1 --steps 1 > gpurun_out/ncu_c2_full4.log 2>&1
ls -la gpurun_out/c2_full4.ncu-rep
