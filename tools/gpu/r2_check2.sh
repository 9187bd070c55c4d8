#!/bin/bash
# round 2: compositor tests, then per-kernel device times of one tile-route step (1080p synthetic, bundled view 0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py tests/test_view_prep.py tests/test_reference_function.py tests/test_views_native.py -x -q -m gpu > gpurun_out/comp.log 2>&1; echo "compositor tests rc=$?"; tail -12 gpurun_out/comp.log
{
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 0 --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
} > gpurun_out/tile_breakdown.log 2>&1
grep -v "cudaLaunch\|cudaStream\|cudaEvent\|Activity\|warn\|profiler" gpurun_out/tile_breakdown.log | head -60
