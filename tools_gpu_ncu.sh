#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { # name op variant halo
  python tools/prof_one.py --op $2 --variant $3 --halo $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_$2 -s 3 -c 1 -f -o gpurun_out/prof_$1 \
      python tools/prof_one.py --op $2 --variant $3 --halo $4 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; cat gpurun_out/plain_$1.log
}
run fwd_final fwd -1 1
run bwd_final bwd -1 1
