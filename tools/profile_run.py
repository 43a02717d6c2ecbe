"""Minimal driver for ncu: plan + one (or a few) MPFA and MPSA assemblies of a bench workload.
    python tools/profile_run.py <workload> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200.fv import scalar_bc_codes, vector_bc_codes  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "tet100k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind, dims, _ = bench.WORKLOADS[w]
g = bench.make_grid(kind, dims)
k, bc, C, vbc = bench.make_params(g)
plan = pb.DevicePlan.for_grid(g)
eta = pb.determine_eta(g)
plan.mpfa_upload(k.values, scalar_bc_codes(bc, g.num_faces), None, eta)
codes, robw = vector_bc_codes(vbc, 3, g.num_faces)
alphas = []
if os.environ.get("PB_BIOT"):
    import numpy as np
    a = np.zeros((3, 3, g.num_cells))
    a[0, 0] = a[1, 1] = a[2, 2] = 0.8
    alphas = [a]
plan.mpsa_upload(C.values, codes, robw, eta, alphas)
for _ in range(reps):
    a = plan.mpfa_assemble()
    b = plan.mpsa_assemble()
print(w, g.num_cells, "cells; mpfa ms", a, "mpsa ms", b)
