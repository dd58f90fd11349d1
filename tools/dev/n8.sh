#!/bin/bash
# the 8-GPU measurements of round 2 (one box, one process per GPU)
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 300 $TR --master-port 29601 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err; echo "A rc=$?"
timeout 400 $TR --master-port 29602 bench.py --gpus $N --workload rmcl-rmat22e32 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/r2_rmcl22_n$N.json 2> gpurun_out/r2_rmcl22_n$N.err; echo "B rc=$?"
timeout 300 $TR --master-port 29603 bench.py --gpus $N --workload stencil256 --rmcl-leg none --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_stencil256_n$N.json 2> gpurun_out/r2_stencil256_n$N.err; echo "C rc=$?"
python - <<PY
import json
for f in ("r2_n$N", "r2_rmcl22_n$N", "r2_stencil256_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["metric"], round(d["value"], 3), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", (d.get("e2e") or {}).get("value"))
        if d.get("rmcl"): print("   rmcl leg:", d["rmcl"].get("value"), d["rmcl"].get("ms_per_step"), d["rmcl"].get("error"))
        if "per_iteration" in d.get("roofline", {}): print("   ", d["roofline"]["per_iteration"]["ms"], d["roofline"]["per_iteration"]["row_tiles"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/r2_rmcl22_n$N.err | cut -c1-300
