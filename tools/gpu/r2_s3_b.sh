#!/bin/bash
# session 3, run B: the radix binning of the tile route — bit-exact tests, then per-kernel device times of both
# binnings on the 1080p view and the bundled view 0
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py -x -q -m gpu -k "binning or pair_list or radix" > gpurun_out/s3_bin_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s3_bin_tests.log
for b in 0 1; do for c2 in -1 0; do
echo "== binning $b c2 $c2"
SPLAT_PROFILE=1 timeout 300 python tools/splat_time.py --route tiles --binning $b --c2 $c2 --steps 6 2>&1 | grep "k_view\|splat step\|elements\|per step" | cut -c1-70,150-215
done; done > gpurun_out/s3_bin_breakdown.log 2>&1
cat gpurun_out/s3_bin_breakdown.log
