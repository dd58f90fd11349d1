#!/bin/bash
# One gpurun call that re-establishes the state of the build on a fresh B200 box (about 3 minutes):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/round_start_gpu.sh'
# Every python process runs under `timeout`: a process that hangs at exit (as a joinable helper
# thread in a static object once did) would otherwise burn the whole limit.
set -u
mkdir -p gpurun_out
echo "== gpu tests"; date +%s
timeout 120 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -8
echo "== bench (default flags)"; date +%s
timeout 150 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n1.json"))
    print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 1),
          "roofline", round(d["roofline"]["frac"], 3), "e2e", d["e2e"].get("value"), "ms", d["e2e"].get("ms_per_step"),
          "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 gpurun_out/bench_n1.err
echo "== streamed product timeline"; date +%s
B200_PROF=1 timeout 60 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err
grep "b200 stream" gpurun_out/bench_prof.err | tail -22 | cut -c1-190
echo "== ncu launch list"; date +%s
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
date +%s
