#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py tests/test_views_native.py tests/test_view_prep.py tests/test_reference_function.py -x -q -m gpu > gpurun_out/comp.log 2>&1; echo "compositor tests rc=$?"; tail -3 gpurun_out/comp.log
python tools/splat_time.py --route tiles --steps 8 2>&1 | grep "per step\|splat step"
python tools/splat_time.py --route tiles --c2 0 --steps 8 2>&1 | grep "per step\|splat step"
