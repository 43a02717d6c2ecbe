#!/bin/bash
# round 2, GPU call 3: fused Krylov + device pool validation, bench line, tet-kernel variants (wide team, adjugate inverse)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_krylov.py tests/test_gpu_parity.py -m gpu -x -q -k "bicgstab or device_side or sharded or split") > gpurun_out/c3_pytest.log 2>&1
tail -3 gpurun_out/c3_pytest.log
for v in base wide adj adjwide; do
  if [ "$v" = base ]; then unset POREB200_LIB; else export POREB200_LIB=$PWD/porepy_b200/libporeb200_$v.so; fi
  echo "== $v"
  python tools/profile_run.py tet1m 3 2>&1 | tail -1
  python tools/profile_run.py cart128 3 2>&1 | tail -1
  if [ "$v" != base ]; then timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_digest_gpu.py -q -m gpu -x -k "golden or seeded or contrast or tet12 or cart16" 2>&1 | tail -1; fi
done > gpurun_out/c3_ab.log 2>&1
unset POREB200_LIB
cat gpurun_out/c3_ab.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c3_bench_tet1m.json 2> gpurun_out/c3_bench_tet1m.err
grep "e2e call" gpurun_out/c3_bench_tet1m.err | cut -c1-420
python - <<'PY'
import json
d = json.load(open("gpurun_out/c3_bench_tet1m.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("krylov", d["krylov"])
print("spmv", {k: (v["ms"], v["frac"]) for k, v in d["spmv"].items()})
print("variants", d["e2e"].get("variants"))
PY
