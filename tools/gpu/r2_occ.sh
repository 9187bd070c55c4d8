#!/bin/bash
mkdir -p gpurun_out
for occ in 3 4 2; do
echo "== GCP_BWD_OCC=$occ"
GCP_BWD_OCC=$occ SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --steps 6 2>&1 | grep "k_view_backward\|k_view_render\|splat step" | cut -c1-70,150-215
GCP_BWD_OCC=$occ SPLAT_PROFILE=1 python tools/splat_time.py --route tiles --c2 0 --steps 6 2>&1 | grep "k_view_backward\|splat step" | cut -c1-70,150-215
done
GCP_BWD_OCC=4 timeout 600 python -m pytest tests/test_compositor.py -x -q -m gpu -k "fixture or ragged or oracle" 2>&1 | tail -2
