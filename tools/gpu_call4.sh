#!/bin/bash
# round 2, GPU call 4 (2 GPUs): device AD tests, sharded strong-scaling bench at N = 2
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_zz_ad_gpu.py -m gpu -x -q) > gpurun_out/c4_pytest_ad.log 2>&1
tail -5 gpurun_out/c4_pytest_ad.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4_bench_n2.json 2> gpurun_out/c4_bench_n2.err
grep -v OpenBLAS gpurun_out/c4_bench_n2.err | tail -12 | cut -c1-500
python - <<'PY'
import json
d = json.load(open("gpurun_out/c4_bench_n2.json"))
print("value", d["value"], "ms", d["ms_per_step"], d["config"]["parallelism"], "halo", d["config"]["halo_cell_overhead"])
print("e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"], d["e2e"]["breakdown"]["stages_s"])
print("krylov", d["krylov"])
PY
