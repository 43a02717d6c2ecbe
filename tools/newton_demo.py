"""Device-resident Newton loop at bench size (porepy_b200/newton.py): nonlinear flow k = k0 exp(beta p) on 998,250
tetrahedra, Jacobian chain + fused BiCGStab on one GPU, no matrix D2H.   python tools/newton_demo.py [workload] [beta]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from porepy_b200 import newton  # noqa: E402
from porepy_b200.sparse import LazyCsr  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "tet1m"
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
kind, dims, _ = bench.WORKLOADS[w]
g = bench.make_grid(kind, dims)
rng = np.random.default_rng(0)
nc = g.num_cells
q = rng.standard_normal((nc, 3, 3))
k0 = (np.einsum("cij,ckj->cik", 0.2 * q, 0.2 * q) + np.eye(3)).reshape(-1)
bf = g.get_all_boundary_faces()
x = g.face_centers[0, bf]
dirf = bf[(x < 1e-10) | (x > 1 - 1e-10)]
dirv = np.where(g.face_centers[0, dirf] < 0.5, 1.0, 0.0)
t0 = time.perf_counter()
prob = newton.NonlinearTpfaFlow(g, k0, beta, dirf, dirv, 0.5 * g.cell_volumes * rng.random(nc))
prob._upload()
setup = time.perf_counter() - t0
d0 = sum(LazyCsr.downloads.values())
t0 = time.perf_counter()
p, hist = newton.solve(prob, tol=1e-9, linear_tol=1e-10, verbose=False)
total = time.perf_counter() - t0
print(json.dumps({"workload": w, "cells": nc, "beta": beta, "setup_s": setup, "newton_s": total,
                  "iterations": len(hist) - 1, "matrix_bytes_to_host": int(sum(LazyCsr.downloads.values()) - d0),
                  "history": hist, "p_min_max": [float(p.min()), float(p.max())]}, indent=1))
