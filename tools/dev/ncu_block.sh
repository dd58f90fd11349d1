#!/bin/bash
# usage: ncu_block.sh <lo> <hi> <tag> <kernel regex>  — ncu --set full of one kernel on one row block
set -u
mkdir -p gpurun_out
export B200_ON_CHIP=1
timeout 100 python tools/hub_block.py $1 $2 > gpurun_out/plain.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o gpurun_out/r2_$3 python tools/hub_block.py $1 $2 > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
