#!/bin/bash
# round 2, GPU call 18 (1 GPU): the GPU tests added this round, bench N = 1, extra timings, final ncu evidence
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_diff_tpfa.py tests/test_geometry.py tests/test_mdg_network.py tests/test_krylov.py -m gpu -q) > gpurun_out/c18_pytest_new.log 2>&1
tail -3 gpurun_out/c18_pytest_new.log | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c18_bench_n1.json 2> gpurun_out/c18_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c18_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   krylov", d.get("krylov"))
print("   cpu", d.get("cpu_baseline"))
PY
tail -3 gpurun_out/c18_bench_n1.err | cut -c1-300
timeout 600 python tools/extra_bench.py tet1m > gpurun_out/c18_extra.json 2> gpurun_out/c18_extra.err; cat gpurun_out/c18_extra.json | head -60; tail -3 gpurun_out/c18_extra.err
bash tools/gpu_final_profile.sh 2>&1 | tail -60
