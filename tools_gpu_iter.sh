#!/usr/bin/env bash
# iteration GPU call: parity tests, variant sweep, bench line
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --sweep --steps 10 > gpurun_out/sweep_c3.json 2> gpurun_out/sweep_c3.txt; echo "sweep rc=$?"
cat gpurun_out/sweep_c3.txt
