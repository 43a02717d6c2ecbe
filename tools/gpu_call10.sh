#!/bin/bash
# round 2, GPU call 10 (2 GPUs): full GPU suite, 2-rank diagnosis at bench size, bench N = 2
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/c10_pytest_all.log 2>&1
tail -4 gpurun_out/c10_pytest_all.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/debug_n2.py tet1m 2>&1 | grep "^\[rank" | cut -c1-600
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c10_bench_n2.json 2> gpurun_out/c10_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c10_bench_n2.json"))
print("N 2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   krylov", d["krylov"])
PY
