"""Device-resident single-phase flow solve: discretize (MPFA, flux terms only) -> A = div @ flux on the
device -> Jacobi-BiCGStab on the device matrix.  No discretization matrix ever leaves HBM; only the
right-hand side and the pressure cross PCIe.   python tools/flow_solve_demo.py [workload]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import krylov as kr  # noqa: E402
from porepy_b200.fv import scalar_bc_codes  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "tet1m"
kind, dims, desc = bench.WORKLOADS[w]
g = bench.make_grid(kind, dims)
k, bc, _, _ = bench.make_params(g)
bf = g.get_all_boundary_faces()
bv = np.zeros(g.num_faces)
bv[bf[g.face_centers[0, bf] < 1e-10]] = 1.0
torch.cuda.synchronize()
t0 = time.perf_counter()
plan = pb.DevicePlan.for_grid(g)
t1 = time.perf_counter()
plan.mpfa_upload(k.values, scalar_bc_codes(bc, g.num_faces), None, pb.determine_eta(g))
ms = plan.mpfa_assemble(True, False, False)
A = plan.mpfa_system()
b = plan.mpfa_rhs(bv)
torch.cuda.synchronize()
t2 = time.perf_counter()
# diagonal for the Jacobi preconditioner: one SpMV-free pass over the device matrix is not exposed yet ->
# take it from a host copy of the (small) diagonal positions
Ah = A.to_scipy()
diag = torch.as_tensor(Ah.diagonal(), dtype=torch.float64, device="cuda")
loc = kr.LocalSystem(0, 1, np.arange(g.num_cells), np.zeros(0, dtype=np.int64), Ah, [0], [np.zeros(0, dtype=np.int64)])
op = kr.DistributedOperator.__new__(kr.DistributedOperator)
op.torch, op.loc, op.device, op.group = torch, loc, torch.device("cuda", 0), None
op.n_own, op.n_ghost = g.num_cells, 0
op.xbuf = torch.zeros(g.num_cells, dtype=torch.float64, device="cuda")
op.send_idx, op.halo_bytes, op._matvec, op.dev_csr = [], 0, None, A
torch.cuda.synchronize()
t3 = time.perf_counter()
x, info = kr.bicgstab(op, torch.as_tensor(b, device="cuda"), tol=1e-8, maxiter=20000, diag_own=diag)
torch.cuda.synchronize()
t4 = time.perf_counter()
p = x.cpu().numpy()
res = float(np.linalg.norm(Ah @ p - b) / np.linalg.norm(b))
print(json.dumps({"workload": desc, "cells": g.num_cells, "plan_s": t1 - t0, "discretize_assemble_s": t2 - t1,
                  "mpfa_kernel_ms": ms, "krylov_s": t4 - t3, "iterations": info["iterations"],
                  "converged": info["converged"], "true_relres": res, "spmv": info["spmv"],
                  "total_s_without_diag_copy": (t2 - t0) + (t4 - t3)}))
