"""Single-GPU diagnosis of the sharded assembly: flux rows of a shard (with / without the active-node mask, device /
host topology plan) against the rows of the whole-mesh discretization."""
import os
import sys

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import shard as sh  # noqa: E402

kind, dims, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "tet10k"]
g = bench.make_grid(kind, dims)
k, bc, C, vbc = bench.make_params(g)
dg = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
bv = np.zeros(g.num_faces)
bfc = g.get_all_boundary_faces()
bv[bfc[g.face_centers[0, bfc] < 1e-10]] = 1.0
dg[pb.PARAMETERS]["flow"]["bc_values"] = bv
mgl = pb.Mpfa("flow")
mgl.discretize(g, dg)
Ag_lazy, bg = mgl.assemble_matrix_rhs(g, dg)
Ag = Ag_lazy.device_csr.to_scipy()
F = sps.csr_matrix(dg[pb.DISCRETIZATION_MATRICES]["flow"]["flux"])
Ag_host = (g.divergence(1) @ F).tocsr()
print("global: device A vs host div@flux", abs(Ag - Ag_host).max(), flush=True)
part = sh.partition_cells(g, 2)
fn = sps.csc_matrix(g.face_nodes)
for r in range(2):
    s = sh.extract_shard(g, part, r)
    n_own = int(s.own_cell.sum())
    lcf = sps.csc_matrix(s.grid.cell_faces)
    own_faces = np.unique(lcf.indices[:lcf.indptr[n_own]])
    for mode in ("active+device", "all+device", "active+host", "active+device+cellmap"):
        if hasattr(s.grid, "_b200_plan"):
            del s.grid._b200_plan
        if "host" in mode:
            os.environ["POREB200_HOST_PLAN"] = "1"
        plan = pb.DevicePlan.for_grid(s.grid)
        os.environ.pop("POREB200_HOST_PLAN", None)
        if mode.startswith("active"):
            plan.set_active_nodes(s.own_node)
        kl = pb.SecondOrderTensor.from_values(s.restrict_cell_array(k.values))
        if "cellmap" in mode:
            plan.set_cell_map(s.cells, g.num_cells)
            kl = k
        dl = pb.initialize_data({}, "flow", {"second_order_tensor": kl, "bc": sh.restrict_scalar_bc(bc, s),
                                             "bc_values": bv[s.faces], "mpfa_eta": pb.determine_eta(g)})
        ml = pb.Mpfa("flow")
        ml.discretize(s.grid, dl)
        a_dev, b_loc = ml.assemble_matrix_rhs_device(s.grid, dl)
        rows = a_dev.to_scipy()[:n_own]
        refA = Ag_host[s.cells[:n_own]][:, s.cells]
        dA = abs(refA - rows)
        badA = np.flatnonzero(np.asarray(dA.max(axis=1).todense()).ravel() > 1e-10)
        print(f"rank {r} {mode:22s}: A rows max err {dA.max():.3e}, bad rows {badA.size}, diag err "
              f"{np.abs(a_dev.diagonal()[:n_own] - Ag_host.diagonal()[s.cells[:n_own]]).max():.3e}, rhs err "
              f"{np.abs(b_loc[:n_own] - bg[s.cells[:n_own]]).max():.3e}", flush=True)
        Fl = sps.csr_matrix(dl[pb.DISCRETIZATION_MATRICES]["flow"]["flux"])
        ref = F[s.faces[own_faces]][:, s.cells]
        d = abs(ref - Fl[own_faces])
        rowerr = np.asarray(d.max(axis=1).todense()).ravel()
        bad = np.flatnonzero(rowerr > 1e-10 * abs(F).max())
        print(f"rank {r} {mode:14s}: own faces {own_faces.size}, bad rows {bad.size}, max err {rowerr.max():.3e}", flush=True)
        if bad.size and mode == "active+device":
            lfn = sps.csc_matrix(s.grid.face_nodes)
            for f in own_faces[bad[:4]]:
                nodes = lfn.indices[lfn.indptr[f]:lfn.indptr[f + 1]]
                cn = (abs(s.grid.face_nodes) @ abs(s.grid.cell_faces)).tocsr()
                print("   face", int(f), "global", int(s.faces[f]), "center", np.round(s.grid.face_centers[:, f], 3).tolist(),
                      "nodes", nodes.tolist(), "own_node", s.own_node[nodes].tolist(),
                      "cells/node", [int(cn[n].nnz) for n in nodes],
                      "glob cells/node", [int((abs(g.face_nodes) @ abs(g.cell_faces)).tocsr()[s.nodes[n]].nnz) for n in nodes], flush=True)
