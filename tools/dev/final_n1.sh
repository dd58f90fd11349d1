#!/bin/bash
# round-2 evidence on one B200: tests, bench (both arms), ncu launch list, ncu full of the top kernel
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== bench"; timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_n1_reference.json 2> gpurun_out/r2_bench_n1_reference.err; echo rc=$?
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 1), "kernel frac", round(r["frac"], 3), "step_frac", round(r["step_frac"], 3),
      "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
m = d["rmcl"]; print("rmcl leg", m.get("value"), m.get("ms_per_step"), m.get("error"), "cpu", (m.get("cpu_baseline") or {}).get("value"), "e2e", (m.get("e2e") or {}).get("value"))
print(json.loads(open("gpurun_out/r2_bench_n1_reference.json").read().strip().splitlines()[-1])["value"])
PY
echo "== ncu launch list"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --rmcl-leg none"
timeout 200 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_rmat20.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "== ncu full of the part kernel"
timeout 280 ncu --set full --clock-control none --import-source on -k regex:k_num_bitmap_part -s 1 -c 1 -f -o gpurun_out/r2_prof_numpart_rmat20 $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
