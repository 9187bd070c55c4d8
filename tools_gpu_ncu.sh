#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { # name op variant halo
  python tools/prof_one.py --op $2 --variant $3 --halo $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_$2 -s 6 -c 2 -f -o gpurun_out/prof_$1 \
      python tools/prof_one.py --op $2 --variant $3 --halo $4 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; cat gpurun_out/plain_$1.log
}
run fwd_tma fwd 4 1
run bwd_tma bwd 3 1
ls -la gpurun_out/*.ncu-rep
