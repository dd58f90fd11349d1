#!/bin/bash
set -u
mkdir -p gpurun_out
export B200_PROF=1 B200_ON_CHIP=1
for blk in "0 1538" "200000 260000" "700000 1048576" "0 1048576"; do echo "== rows $blk"; timeout 100 python tools/hub_block.py $blk 2>&1 | grep -E "k_num_units|rows \[" | tail -2; done
unset B200_PROF B200_ON_CHIP
true &&
true
tail -2 gpurun_out/ncu_mid.log
