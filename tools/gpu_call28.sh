#!/bin/bash
# round 2, final validation on one GPU: smoke(), full GPU suite, bench N = 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/c28_pytest_all.log 2>&1
tail -5 gpurun_out/c28_pytest_all.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/c28_pytest_all.log | head
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c28_bench_n1.json 2> gpurun_out/c28_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c28_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres")})
print("   mech solve", {q: (k.get("mechanics") or {}).get(q) for q in ("iterations", "converged", "seconds", "error")})
print("   spmv", {q: (round(v["ms"], 4), round(v["frac"], 3)) for q, v in (d.get("spmv") or {}).items()})
print("   roofline", {q: d["roofline"].get(q) for q in ("achieved", "frac", "traffic", "fp64_frac_of_measured_dmma")})
print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "e2e bytes", d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], "launches", d["gpu_launches"])
PY
grep "e2e call" gpurun_out/c28_bench_n1.err | sed "s/{.stages_s.*device_alloc_outside_pool/ alloc/" | cut -c1-200
