#!/usr/bin/env bash
# first GPU call: smoke, parity tests, variant sweep, bench line, reference fixture
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tests/golden/make_ref_fixture.py > gpurun_out/fixture.log 2>&1; echo "fixture rc=$?"
timeout 600 python bench.py --sweep --steps 10 > gpurun_out/sweep_c3.json 2> gpurun_out/sweep_c3.txt; echo "sweep rc=$?"
cat gpurun_out/sweep_c3.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
