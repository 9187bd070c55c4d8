#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/splat_time.py --steps 3 > gpurun_out/plain_splat.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_splat.csv \
  python tools/splat_time.py --steps 3 > gpurun_out/ncu_splat.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/plain_splat.log
