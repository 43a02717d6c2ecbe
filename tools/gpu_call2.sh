#!/bin/bash
# round 2, GPU call 2: full GPU suite on the lazy / sharded code, bench line, ncu captures of the two MPSA kernels
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/c2_pytest.log 2>&1
tail -4 gpurun_out/c2_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c2_bench_tet1m.json 2> gpurun_out/c2_bench_tet1m.err
tail -c 1500 gpurun_out/c2_bench_tet1m.json; tail -5 gpurun_out/c2_bench_tet1m.err
bash tools/ncu_capture.sh c2_mpsa_cart64 'mpsa_kernel.*TileGJ<.int.2,..int.3,..int.8' cart64 1 > gpurun_out/c2_ncu1.log 2>&1
bash tools/ncu_capture.sh c2_mpsa_tet100k 'mpsa_kernel.*TileGJ<.int.7,..int.2,..int.24' tet100k 1 > gpurun_out/c2_ncu2.log 2>&1
tail -3 gpurun_out/c2_ncu1.log gpurun_out/c2_ncu2.log
du -sh gpurun_out
