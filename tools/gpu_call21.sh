#!/bin/bash
# round 2, GPU call 21 (8 GPUs): BASELINE config[2] size -- Cartesian 128^3 (2,097,152 cells) sharded over 8 GPUs, then 1 GPU
mkdir -p gpurun_out
for n in 8 1; do
if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"; fi
timeout 900 $L bench.py --gpus $n --steps 5 --warmup 3 --workload cart128 > gpurun_out/c21_bench_cart128_n$n.json 2> gpurun_out/c21_bench_cart128_n$n.err
python - <<PY
import json
d = json.load(open("gpurun_out/c21_bench_cart128_n$n.json"))
print("cart128 N $n value", d["value"], "ms", d["ms_per_step"], "mpfa/mpsa", d["config"]["ms_mpfa"], d["config"]["ms_mpsa"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
k = d.get("krylov") or {}
print("   flow solve", {q: k.get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres")})
print("   mech solve", {q: (k.get("mechanics") or {}).get(q) for q in ("iterations", "converged", "seconds", "ms_per_iteration", "true_relres", "error")})
print("   spmv", {q: (round(v["ms"], 4), round(v["frac"], 3)) for q, v in (d.get("spmv") or {}).items()})
PY
done
export CUDA_VISIBLE_DEVICES=0
timeout 300 python tools/plan_stage_profile.py 2>&1 | grep "^call" | cut -c1-300
