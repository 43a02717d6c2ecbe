#!/bin/bash
# round 2, GPU call 16 (2 GPUs): fused BiCGStab after the SpMV warp-uniform fix, plan phases, bench N = 2 and N = 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/debug_n2.py tet1m > gpurun_out/c16_debug.log 2>&1
grep "^\[rank" gpurun_out/c16_debug.log | cut -c1-700
POREB200_PLAN_TIMING=1 timeout 300 python tools/plan_timing.py > gpurun_out/c16_plan.log 2>&1; tail -45 gpurun_out/c16_plan.log
for n in 2 1; do
if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514"; fi
timeout 900 $L bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/c16_bench_n$n.json 2> gpurun_out/c16_bench_n$n.err
python - <<PY
import json
d = json.load(open("gpurun_out/c16_bench_n$n.json"))
print("N $n value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   krylov", d.get("krylov"))
PY
done
