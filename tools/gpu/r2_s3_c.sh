#!/bin/bash
# session 3, run C: radix binning — bit-exact tests, per-kernel device times (radix only) on the 1080p view and bundled view 0
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py -x -q -m gpu -k "binning or pair_list or radix" > gpurun_out/s3_bin_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s3_bin_tests.log
for c2 in -1 0; do
echo "== binning 1 c2 $c2"
SPLAT_PROFILE=1 timeout 300 python tools/splat_time.py --route tiles --binning 1 --c2 $c2 --steps 6 2>&1 | grep "k_view\|splat step\|elements\|per step" | cut -c1-70,150-215
done > gpurun_out/s3_bin_breakdown.log 2>&1
grep -v "k_view_\(backward\|reduce\|cnt\|combine\|render\)" gpurun_out/s3_bin_breakdown.log
