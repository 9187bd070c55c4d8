#!/usr/bin/env bash
# iteration GPU call: parity tests, variant sweep
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --sweep --steps 10 > gpurun_out/sweep_c3.json 2> gpurun_out/sweep_c3.txt; echo "sweep rc=$?"
grep -E "blk|copy" gpurun_out/sweep_c3.txt | grep -v "halo=0"
