#!/bin/bash
set -u
N=$1; FULL=${2:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29603 bench.py --gpus $N --workload stencil256 --rmcl-leg none --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_stencil256_n$N.json 2> gpurun_out/r2_stencil256_n$N.err; echo "stencil rc=$?"
if [ $FULL = 1 ]; then timeout 400 $TR --master-port 29601 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err; echo "default rc=$?"; fi
python - <<PY
import json
for f in ("r2_stencil256_n$N", "r2_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "ranks", d["config"].get("rank_ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"))
        if d.get("rmcl"): print("   rmcl leg:", d["rmcl"].get("value"), d["rmcl"].get("ms_per_step"), d["rmcl"]["roofline"].get("loops_ms_rank0"), d["rmcl"]["roofline"]["per_iteration"]["ms"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
