#!/bin/bash
# compositor tests, then per-kernel device times of one step through both routes (1080p synthetic and bundled view 1)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_compositor.py tests/test_view_prep.py -x -q -m gpu 2>&1 | tail -15
for r in tiles lists; do
SPLAT_PROFILE=1 python tools/splat_time.py --route $r --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
SPLAT_PROFILE=1 python tools/splat_time.py --route $r --c2 1 --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
done
} > gpurun_out/tile_check.log 2>&1
tail -5 gpurun_out/tile_check.log
