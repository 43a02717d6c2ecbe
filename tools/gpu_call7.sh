#!/bin/bash
# round 2, GPU call 7 (2 GPUs): rotated solver loop (parity + timing, it is the default build now; PB_NO_ROT lib as base),
# 2-rank diagnosis of the sharded flow solve at bench size
mkdir -p gpurun_out
for v in norot base; do
  if [ "$v" = base ]; then unset POREB200_LIB; else export POREB200_LIB=$PWD/porepy_b200/libporeb200_$v.so; fi
  echo "== $v"
  python tools/profile_run.py tet1m 3 2>&1 | tail -1
  python tools/profile_run.py cart128 3 2>&1 | tail -1
done 2>&1 | grep -v OpenBLAS
unset POREB200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_digest_gpu.py tests/test_zz_tpfa_upwind_gpu.py -q -m gpu -x 2>&1 | grep -v OpenBLAS | tail -3
for w in tet100k tet1m; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/debug_n2.py $w 2>&1 | grep "^\[rank" | cut -c1-700
done
