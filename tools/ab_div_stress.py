"""A/B of the two forms of A = div_nd @ stress on the device (csrc/api.cu): scatter with atomics (POREB200_DIV_SCATTER=1)
against the gather form (default when the plan holds the cell -> face lists).  Prints wall milliseconds per call (device
synchronised) and the checksums of both results.   python tools/ab_div_stress.py [tet1m|cart128|tet100k]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200.fv import vector_bc_codes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tet1m"
kind, dims, _ = bench.WORKLOADS[name]
g = bench.make_grid(kind, dims)
k, bc, C, vbc = bench.make_params(g)[:4]
plan = pb.DevicePlan.for_grid(g)
codes, robw = vector_bc_codes(vbc, 3, g.num_faces)
plan.mpsa_upload(C.values, codes, robw, pb.determine_eta(g))
plan.mpsa_assemble()
res = {}
for mode in ("scatter", "gather", "scatter", "gather"):
    if mode == "scatter":
        os.environ["POREB200_DIV_SCATTER"] = "1"
    else:
        os.environ.pop("POREB200_DIV_SCATTER", None)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = plan.mpsa_system()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        chk = a.checksum()
        del a
    res.setdefault(mode, []).append((min(ts), chk))
    print(f"{name} {mode:8s} ms per pb_mpsa_system call {['%.2f' % t for t in ts]}  checksum {chk}", flush=True)
s, q = res["scatter"][0][1], res["gather"][0][1]
print("relative difference of the checksums", abs(s[0] - q[0]) / max(abs(s[0]), 1e-300), abs(s[1] - q[1]) / abs(s[1]))
