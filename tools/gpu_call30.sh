#!/bin/bash
# round 2, call 30 (final tree): smoke(), the whole GPU suite, bench N = 1 (with the mixed-dimensional extra), PCIe probe
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 1500 python -m pytest tests -m gpu -q -s) > gpurun_out/c30_pytest_all.log 2>&1
tail -5 gpurun_out/c30_pytest_all.log | cut -c1-300; grep -n "^FAILED\|^ERROR\|against the oracle" gpurun_out/c30_pytest_all.log | head
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c30_bench_n1.json 2> gpurun_out/c30_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c30_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds")})
print("   mech solve", {q: (k.get("mechanics") or {}).get(q) for q in ("iterations", "converged", "seconds", "error")})
print("   roofline", {q: d["roofline"].get(q) for q in ("achieved", "frac", "traffic", "fp64_frac_of_measured_dmma")})
print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "launches", d["gpu_launches"])
print("   md", json.dumps(d.get("md_network"))[:1800])
PY
tail -2 gpurun_out/c30_bench_n1.err | cut -c1-300
python - <<PY
import torch, time
a = torch.empty(1 << 27, dtype=torch.float64).pin_memory()      # 1 GiB
b = torch.empty_like(a, device="cuda")
for name, src, dst in (("H2D", a, b), ("D2H", b, a)):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"PCIe {name} 1 GiB pinned: {1.073741824 / dt:.1f} GB/s")
PY
