#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py -x -q -m gpu 2>&1 | tail -5
python tools/splat_time.py --c2 0 --steps 3 && python tools/splat_time.py --steps 5 && python tools/splat_time.py --c2 1 --steps 3 && \
ncu --set full --clock-control none --import-source on -k regex:'k_splat|k_place' --launch-skip 20 -c 10 -f -o gpurun_out/c2_full python tools/splat_time.py --c2 1 --steps 1 > gpurun_out/ncu_c2_full.log 2>&1
ls -la gpurun_out/c2_full.ncu-rep
