#!/bin/bash
# round 2, GPU call 6 (2 GPUs): bench N = 1 and N = 2 with the device topology plan, graph Krylov, cell map
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_krylov.py tests/test_zz_multigpu.py -m gpu -x -q -k "sharded or bicgstab or two_gpu or device_side" 2>&1 | grep -v OpenBLAS | tail -4
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f"gpurun_out/c6_bench_n{n}.json"))
    except Exception as e:
        print("N", n, "no json:", e); continue
    print("N", n, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
    print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
    print("   krylov", d["krylov"])
PY
grep -v OpenBLAS gpurun_out/c6_bench_n2.err | tail -5 | cut -c1-300
