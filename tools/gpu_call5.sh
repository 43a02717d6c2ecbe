#!/bin/bash
# round 2, GPU call 5 (2 GPUs): device topology plan vs host plan, 2-GPU NCCL test, then the full GPU suite
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_zz_device_plan_gpu.py tests/test_zz_multigpu.py -m gpu -x -q -s) > gpurun_out/c5_pytest_a.log 2>&1
grep -v OpenBLAS gpurun_out/c5_pytest_a.log | tail -25 | cut -c1-1500
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/c5_pytest_all.log 2>&1
tail -6 gpurun_out/c5_pytest_all.log | cut -c1-600
POREB200_PLAN_TIMING=1 python tools/profile_run.py tet1m 1 2>&1 | grep -v OpenBLAS | tail -12
