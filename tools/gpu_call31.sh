#!/bin/bash
# round 2, call 31: bench N = 1 on the final tree (boundary-only basis test, merged interface blocks of the md assembly) + the md GPU tests
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zzz_md_gpu.py -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/c31_bench_n1.json 2> gpurun_out/c31_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c31_bench_n1.json"))
print("N 1 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_step"])
print("   stages", {k: round(v, 4) for k, v in d["e2e"]["breakdown"]["stages_s"].items()})
print("   mpsa", d["e2e"]["breakdown"]["mpsa"])
m = d.get("md_network") or {}
print("   md calls", m.get("calls"), "cells/s", m.get("cells_per_s"), "ad", m.get("assemble_ad_s"), "err", m.get("error"))
print("   md solve", m.get("solve"))
PY
tail -2 gpurun_out/c31_bench_n1.err | cut -c1-300
