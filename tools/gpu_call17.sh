#!/bin/bash
# round 2, GPU call 17 (1 GPU): full GPU suite, plan phases, bench N = 1 (with the mechanics solve)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/c17_pytest_all.log 2>&1
tail -6 gpurun_out/c17_pytest_all.log | cut -c1-600
POREB200_PLAN_TIMING=1 timeout 300 python tools/plan_timing.py > gpurun_out/c17_plan.log 2>&1; tail -40 gpurun_out/c17_plan.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c17_bench_n1.json 2> gpurun_out/c17_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c17_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   krylov", d.get("krylov"))
print("   cpu", d.get("cpu_baseline"))
PY
tail -5 gpurun_out/c17_bench_n1.err | cut -c1-400
