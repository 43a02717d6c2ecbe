"""Timings of the "next rows" at bench size on one GPU (998,250 tets): Grid.compute_geometry on the device, the fused
differentiable-TPFA evaluation, TPFA / upwinding, the native shard extraction.   python tools/extra_bench.py [workload]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import shard as sh  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "tet1m"
kind, dims, _ = bench.WORKLOADS[w]
g = bench.make_grid(kind, dims)
nc, nf = g.num_cells, g.num_faces
out = {"workload": w, "cells": nc, "faces": nf}


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), r


# ---- compute_geometry: device vs the generator's own NumPy geometry (same arrays)
ref = [np.array(getattr(g, k)) for k in ("face_normals", "face_centers", "face_areas", "cell_centers", "cell_volumes")]
pb.compute_geometry(g, assign=False)
t, got = best(lambda: pb.compute_geometry(g, assign=False))
err = max(float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) for a, b in zip(got, ref))
bytes_alg = 8 * (3 * g.num_nodes + 7 * nf + 4 * nc) + 4 * (g.face_nodes.nnz + g.cell_faces.nnz)
out["compute_geometry"] = {"host_to_host_s": t, "kernel_ms": pb.compute_geometry.last_kernel_ms,
                           "cells_per_s_host_to_host": nc / t, "max_rel_diff_vs_generator_geometry": err,
                           "algorithmic_GBps_kernels": bytes_alg / (pb.compute_geometry.last_kernel_ms * 1e-3) / 1e9}
# ---- differentiable TPFA
rng = np.random.default_rng(0)
q = rng.standard_normal((nc, 3, 3))
k_c = (np.einsum("cij,ckj->cik", q, q) + 0.5 * np.eye(3)).reshape(-1)
dt = pb.DifferentiableTpfa()
dt.transmissibility(g, k_c)
t, (T, jac, t_hf) = best(lambda: dt.transmissibility(g, k_c))
out["differentiable_tpfa"] = {"host_to_host_s": t, "faces_per_s": nf / t, "jacobian_nnz": int(jac.nnz),
                              "homogeneity_residual": float(np.abs(jac @ k_c - T).max() / np.abs(T).max())}
# ---- TPFA + upwind through the operator classes
k, bc, C, vbc = bench.make_params(g)
d = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
tp = pb.Tpfa("flow")
tp.discretize(g, d)
t, _ = best(lambda: tp.discretize(g, d))
out["tpfa_discretize_s"] = t
du = pb.initialize_data({}, "transport", {"bc": bc, "darcy_flux": rng.standard_normal(nf)})
up = pb.Upwind("transport")
up.discretize(g, du)
t, _ = best(lambda: up.discretize(g, du))
out["upwind_discretize_s"] = t
# ---- native shard extraction (host)
for parts in (2, 8):
    part = sh.partition_cells(g, parts)
    sh.extract_shard(g, part, 0)
    t, s = best(lambda: sh.extract_shard(g, part, parts - 1))
    tn, _ = best(lambda: sh.extract_shard_numpy(g, part, parts - 1), reps=1)
    out[f"shard_extraction_{parts}_parts"] = {"native_s": t, "numpy_s": tn, "cells_incl_halo": int(s.cells.size),
                                              "own_cells": int(s.own_cell.sum())}
print(json.dumps(out, indent=1))
