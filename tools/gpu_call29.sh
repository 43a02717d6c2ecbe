#!/bin/bash
# round 2, call 29: mixed-dimensional GPU tests, the div_nd @ stress gather kernel (tests + A/B), bench N = 1 with the md extra
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_zzz_md_gpu.py tests/test_gpu_parity.py tests/test_zz_ad_gpu.py tests/test_krylov.py -m gpu -q) > gpurun_out/c29_pytest.log 2>&1
tail -4 gpurun_out/c29_pytest.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/c29_pytest.log | head
timeout 300 python tools/ab_div_stress.py tet1m > gpurun_out/c29_ab_div_stress.log 2>&1; tail -6 gpurun_out/c29_ab_div_stress.log | cut -c1-250
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c29_bench_n1.json 2> gpurun_out/c29_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c29_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds")})
print("   mech solve", {q: (k.get("mechanics") or {}).get(q) for q in ("iterations", "converged", "seconds", "error")})
print("   md", json.dumps(d.get("md_network"))[:1500])
PY
tail -3 gpurun_out/c29_bench_n1.err | cut -c1-300
