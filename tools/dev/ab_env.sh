#!/bin/bash
# usage: ab_env.sh "VAR=1" [more settings ...] — bench headline with and without the settings
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --rmcl-leg none > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/ab_$tag.json")); r = d["roofline"]
print("$tag", "ms/step", round(d["ms_per_step"], 2), "GF/s", round(d["value"], 1), "step_frac", round(r["step_frac"], 3), {k: round(v, 1) for k, v in list(r["kernels_ms"].items())[:3]}, r["phases_ms"])
PY
}
run base B200_NOOP=1
run test "$@"
