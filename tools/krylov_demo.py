"""Distributed BiCGStab demo on the MPFA flow matrix (one process per GPU):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/krylov_demo.py [n]
Every rank discretizes the (small) global grid on its GPU, keeps the rows of its slab and solves
A p = b with halo exchange of ghost cells (NCCL point-to-point) and all-reduced dot products."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import _lib, krylov as kr, shard as sh  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
_lib.check(_lib.load().pb_set_device(local))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
g = pb.cart_grid_3d([n, n, n], perturb=0.2, seed=1)
rng = np.random.default_rng(0)
nc = g.num_cells
k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                         0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
bf = g.get_all_boundary_faces()
x = g.face_centers[0, bf]
bc = pb.BoundaryCondition(g, bf[(x < 1e-10) | (x > 1 - 1e-10)], "dir")
bv = np.zeros(g.num_faces)
bv[bf[x < 1e-10]] = 1.0
data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc, "bc_values": bv})
d = pb.Mpfa("flow")
d.discretize(g, data)
A, b = d.assemble_matrix_rhs(g, data)
owner = sh.partition_cells(g, world)
torch.cuda.synchronize()
t0 = time.perf_counter()
xs, owned, info = kr.solve(A, b, owner=owner, tol=1e-10, maxiter=5000)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = None
if world > 1:
    out = [None] * world if rank == 0 else None
    dist.gather_object((owned, xs.cpu().numpy()), out, dst=0)
    if rank == 0:
        full = np.zeros(nc)
        for o, xv in out:
            full[o] = xv
        res = float(np.linalg.norm(A @ full - b) / np.linalg.norm(b))
else:
    res = float(np.linalg.norm(A @ xs.cpu().numpy() - b) / np.linalg.norm(b))
if rank == 0:
    print(json.dumps({"cells": nc, "ranks": world, "iterations": info["iterations"], "converged": info["converged"],
                      "true_relres": res, "seconds": dt, "spmv": info["spmv"], "allreduce": info["allreduce"],
                      "halo_bytes_per_spmv_rank0": info["halo_bytes_per_spmv"]}))
if world > 1:
    dist.destroy_process_group()
