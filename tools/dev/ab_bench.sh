#!/bin/bash
# A/B of the numeric pass on the headline workload: ranges (default) vs part kernel only.
set -u
mkdir -p gpurun_out
for tag in default noranges; do
  if [ $tag = noranges ]; then export B200_NO_RANGES=1; else unset B200_NO_RANGES; fi
  timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$tag.json"))
r = d["roofline"]
print("$tag", "ms/step", round(d["ms_per_step"], 2), "GF/s", round(d["value"], 1), "step_frac", round(r["step_frac"], 3), r["kernels_ms"], r["phases_ms"])
PY
done
unset B200_NO_RANGES
for blk in "0 1538" "200000 260000" "700000 1048576"; do timeout 100 python tools/hub_block.py $blk 2>&1 | tail -1; done
