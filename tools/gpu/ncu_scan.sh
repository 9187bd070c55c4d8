#!/usr/bin/env bash
# ncu --set full of ONE default scan kernel per call:  bash tools/gpu/ncu_scan.sh fwd|bwd
set -u
mkdir -p gpurun_out
op=${1:-bwd}
python tools/prof_one.py --op $op --variant -1 --halo 1 > gpurun_out/plain_${op}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_$op -s 3 -c 1 -f -o gpurun_out/prof_${op}_final \
    python tools/prof_one.py --op $op --variant -1 --halo 1 > gpurun_out/ncu_${op}.log 2>&1
echo "$op rc=$?"; cat gpurun_out/plain_${op}.log
