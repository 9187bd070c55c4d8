#!/bin/bash
# round 2: alignment-peel tests, then per-kernel device times of one tile-route step (1080p synthetic, bundled view 0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alignment_peel or two_streams" > gpurun_out/peel.log 2>&1; echo "peel rc=$?"; tail -5 gpurun_out/peel.log
{
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 0 --steps 6 2>&1 | grep -v "^-----\|aten::\|Memset\|autograd" | cut -c1-70,150-215
} > gpurun_out/tile_breakdown.log 2>&1
cat gpurun_out/tile_breakdown.log | head -80
