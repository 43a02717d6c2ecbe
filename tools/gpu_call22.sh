#!/bin/bash
# round 2, GPU call 22 (1 GPU): full GPU suite on the final tree, Newton demo, bench N = 1
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/c22_pytest_all.log 2>&1
tail -6 gpurun_out/c22_pytest_all.log | cut -c1-400
grep -n "^FAILED\|^ERROR" gpurun_out/c22_pytest_all.log | head
timeout 600 python tools/newton_demo.py tet100k 0.5 > gpurun_out/c22_newton_tet100k.json 2> gpurun_out/c22_newton.err; python - <<'PY'
import json
for f in ("gpurun_out/c22_newton_tet100k.json",):
    try:
        d = json.load(open(f))
        print(f, "cells", d["cells"], "newton_s", round(d["newton_s"], 3), "its", d["iterations"], "matrix bytes to host", d["matrix_bytes_to_host"])
        for h in d["history"]: print("   ", {k: (round(v, 4) if isinstance(v, float) and k != "residual" else v) for k, v in h.items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/c22_newton.err | cut -c1-300
timeout 600 python tools/newton_demo.py tet1m 0.5 > gpurun_out/c22_newton_tet1m.json 2>> gpurun_out/c22_newton.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c22_newton_tet1m.json"))
    print("tet1m newton_s", round(d["newton_s"], 3), "setup_s", round(d["setup_s"], 3), "its", d["iterations"], "matrix bytes to host", d["matrix_bytes_to_host"])
    for h in d["history"]: print("   ", {k: (round(v, 4) if isinstance(v, float) and k != "residual" else v) for k, v in h.items()})
except Exception as e:
    print("tet1m failed", e)
PY
tail -3 gpurun_out/c22_newton.err | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c22_bench_n1.json 2> gpurun_out/c22_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c22_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres")})
print("   spmv", {q: (round(v["ms"], 4), round(v["frac"], 3)) for q, v in (d.get("spmv") or {}).items()})
print("   roofline traffic", d["roofline"].get("traffic"), "frac", d["roofline"]["frac"])
PY
grep "e2e call" gpurun_out/c22_bench_n1.err | cut -c1-200
