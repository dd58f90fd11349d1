#!/bin/bash
# usage: ncu_block.sh <lo> <hi> <tag>   — ncu --set full of k_num_items on one row block
set -u
mkdir -p gpurun_out
timeout 100 python tools/hub_block.py $1 $2 > gpurun_out/plain.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:k_num_items -c 1 -f -o gpurun_out/r2_items_$3 python tools/hub_block.py $1 $2 > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
