#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py tests/test_view_prep.py -x -q -m gpu 2>&1 | tail -3
SPLAT_PROFILE=1 python tools/splat_time.py --c2 1 --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
SPLAT_PROFILE=1 python tools/splat_time.py --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
python tools/splat_time.py --c2 1 --steps 1 && ncu --set full --clock-control none --import-source on -k regex:'k_splat|k_place' --launch-skip 26 -c 13 -f -o gpurun_out/c2_full2 python tools/splat_time.py --c2 1 --steps 1 > gpurun_out/ncu_c2_full2.log 2>&1
ls -la gpurun_out/c2_full2.ncu-rep
