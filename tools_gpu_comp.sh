#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py tests/test_abi.py -x -q -m gpu > gpurun_out/pytest_comp.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_comp.log
timeout 600 python tools/splat_time.py > gpurun_out/splat_time.log 2>&1; echo "splat rc=$?"; tail -4 gpurun_out/splat_time.log
