#!/bin/bash
# session 3, run A: the new maximum-size / chunking / deep-view tests, then the small-n crossover sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_chunking.py tests/test_compositor.py -x -q -m gpu -k "chunk or max_elements or 2_31" > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s3_tests.log
timeout 600 python tools/small_n_sweep.py > gpurun_out/s3_small_n.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/s3_small_n.txt
