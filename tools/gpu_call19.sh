#!/bin/bash
# round 2, GPU call 19 (2 GPUs): the 2-GPU test, fused BiCGStab on two ranks after the SpMV fix, bench N = 2, the
# reference arm under torchrun; then on GPU 0: Krylov micro timings, extra timings, ncu captures of the dominant kernels
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_zz_multigpu.py -m gpu -q) > gpurun_out/c19_pytest_mgpu.log 2>&1; tail -2 gpurun_out/c19_pytest_mgpu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/debug_n2.py tet1m > gpurun_out/c19_debug.log 2>&1
grep "^\[rank" gpurun_out/c19_debug.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c19_bench_n2.json 2> gpurun_out/c19_bench_n2.err
python - <<PY
import json
d = json.load(open("gpurun_out/c19_bench_n2.json"))
print("N 2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   alloc", d["e2e"]["breakdown"].get("device_alloc_outside_pool"))
print("   krylov", d.get("krylov"))
PY
grep "e2e call" gpurun_out/c19_bench_n2.err | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2> gpurun_out/c19_ref.err | cut -c1-700
export CUDA_VISIBLE_DEVICES=0
timeout 600 python tools/krylov_micro.py tet1m 2>&1 | grep -v Warning | tail -20
timeout 600 python tools/extra_bench.py tet1m > gpurun_out/c19_extra.json 2> gpurun_out/c19_extra.err; cat gpurun_out/c19_extra.json | head -60; tail -3 gpurun_out/c19_extra.err
bash tools/gpu_final_profile.sh 2>&1 | tail -45
