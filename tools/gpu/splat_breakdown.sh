#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_compositor.py tests/test_view_prep.py -x -q -m gpu 2>&1 | tail -3
SPLAT_PROFILE=1 python tools/splat_time.py --c2 1 --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
SPLAT_PROFILE=1 python tools/splat_time.py --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
} > gpurun_out/c2e.log 2>&1
