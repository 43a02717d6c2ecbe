#!/bin/bash
# round 2, GPU call 1: full GPU test suite, solver-variant A/B, bench line, host profile of the e2e call, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_smi.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/c1_pytest.log 2>&1
tail -5 gpurun_out/c1_pytest.log
for v in base onebar redux both fine; do
  if [ "$v" = base ]; then unset POREB200_LIB; else export POREB200_LIB=$PWD/porepy_b200/libporeb200_$v.so; fi
  echo "== $v"
  python tools/profile_run.py tet1m 3 2>&1 | tail -1
  python tools/profile_run.py cart128 3 2>&1 | tail -1
  if [ "$v" != base ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or seeded" 2>&1 | tail -1; fi
done > gpurun_out/c1_ab.log 2>&1
unset POREB200_LIB
cat gpurun_out/c1_ab.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_tet1m.json 2> gpurun_out/c1_bench_tet1m.err
tail -c 600 gpurun_out/c1_bench_tet1m.json
timeout 600 python tools/e2e_profile.py tet1m > gpurun_out/c1_e2e_profile.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c1_launches_tet1m.csv \
    python tools/profile_run.py tet1m 1 > /dev/null 2>&1
bash tools/ncu_capture.sh c1_mpsa_cart64 'mpsa_kernel<3,.pb::TileGJ<2,.3,.8' cart64 1 > gpurun_out/c1_ncu1.log 2>&1
bash tools/ncu_capture.sh c1_mpsa_tet100k 'mpsa_kernel<3,.pb::TileGJ<7,.2,.24' tet100k 1 > gpurun_out/c1_ncu2.log 2>&1
rm -f gpurun_out/*_source.csv.tmp
du -sh gpurun_out
