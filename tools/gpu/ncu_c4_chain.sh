#!/usr/bin/env bash
# ncu --set full of the blocked backward on C4 (4K + 512 deep segments): ticketed vs chained tile ranges
set -u
mkdir -p gpurun_out
for c in 0 1; do
python tools/prof_one.py --op bwd --workload c4 --chain $c > gpurun_out/plain_c4_chain$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bwd -s 3 -c 1 -f -o gpurun_out/prof_c4_bwd_chain$c \
    python tools/prof_one.py --op bwd --workload c4 --chain $c > gpurun_out/ncu_c4_chain$c.log 2>&1
echo "chain=$c rc=$?"; cat gpurun_out/plain_c4_chain$c.log
done
