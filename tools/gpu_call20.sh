#!/bin/bash
# round 2, GPU call 20 (N GPUs given as $1): bench at N and at N/2 ... on the same box
mkdir -p gpurun_out
for n in "$@"; do
if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
timeout 900 $L bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/c20_bench_n$n.json 2> gpurun_out/c20_bench_n$n.err
python - <<PY
import json
d = json.load(open("gpurun_out/c20_bench_n$n.json"))
print("N $n value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres")})
print("   mech solve", {q: (k.get("mechanics") or {}).get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres", "error")})
print("   spmv", {q: (round(v["ms"], 4), round(v["frac"], 3)) for q, v in (d.get("spmv") or {}).items()})
print("   config", d["config"].get("parallelism"), d["config"].get("halo_cell_overhead"))
PY
grep "e2e call" gpurun_out/c20_bench_n$n.err | grep "rank 0" | cut -c1-250
done
