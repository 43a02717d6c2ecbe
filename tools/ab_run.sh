#!/bin/bash
# usage: tools/ab_run.sh <variant> ...   (GPU box) -- times alternative builds porepy_b200/libporeb200_<variant>.so
# against the default library on the two bench workloads and runs the GPU parity tests on every variant.
for v in base "$@"; do
  if [ "$v" = base ]; then unset POREB200_LIB; else export POREB200_LIB=$PWD/porepy_b200/libporeb200_$v.so; fi
  echo "== $v"
  python tools/profile_run.py tet1m 3 2>&1 | tail -1
  python tools/profile_run.py cart128 3 2>&1 | tail -1
  if [ "$v" != base ]; then python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2; fi
done
