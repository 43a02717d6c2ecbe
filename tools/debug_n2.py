"""2-rank diagnosis at bench size: distributed SpMV and rhs of the sharded flow system against the whole-mesh system
assembled on the same GPU.   torchrun --nproc-per-node 2 tools/debug_n2.py tet1m"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import porepy_b200 as pb  # noqa: E402
from porepy_b200 import _lib  # noqa: E402
from porepy_b200 import krylov as kr  # noqa: E402
from porepy_b200 import shard as sh  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
_lib.check(_lib.load().pb_set_device(local))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kind, dims, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "tet100k"]
g = bench.make_grid(kind, dims)
k, bc, C, vbc = bench.make_params(g)
bv = np.zeros(g.num_faces)
bf = g.get_all_boundary_faces()
bv[bf[g.face_centers[0, bf] < 1e-10]] = 1.0
dg = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc, "bc_values": bv})
mg = pb.Mpfa("flow")
mg.discretize(g, dg)
Ag, bg = mg.assemble_matrix_rhs(g, dg)
Agd = Ag.device_csr
part = sh.partition_cells(g, world)
s = sh.extract_shard(g, part, rank)
n_own = int(s.own_cell.sum())
pl = pb.DevicePlan.for_grid(s.grid)
pl.set_active_nodes(s.own_node)
pl.set_cell_map(s.cells, g.num_cells)
dl = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": sh.restrict_scalar_bc(bc, s), "bc_values": bv[s.faces],
                                     "mpfa_eta": pb.determine_eta(g)})
ml = pb.Mpfa("flow")
ml.discretize(s.grid, dl)
a_dev, b_loc = ml.assemble_matrix_rhs_device(s.grid, dl)
diag = a_dev.diagonal()[:n_own]
a_dev.truncate_rows(n_own)
loc = kr.local_system_from_shard(s, part, a_dev)
op = kr.DistributedOperator(loc, torch.device("cuda", local))
xg = np.sin(0.37 * np.arange(g.num_cells)) + 0.1
yg = (Agd @ torch.as_tensor(xg, device="cuda")).cpu().numpy()
y = op.matvec(torch.as_tensor(xg[s.cells[:n_own]], device="cuda")).cpu().numpy()
err = np.abs(y - yg[s.cells[:n_own]])
bad = np.flatnonzero(err > 1e-9 * np.abs(yg).max())
dgl = Agd.diagonal()
print(f"[rank {rank}] n_own {n_own} ghosts {loc.ghosts.size}  spmv max err {err.max():.3e} rel {err.max() / np.abs(yg).max():.3e} "
      f"bad rows {bad.size}  rhs err {np.abs(b_loc[:n_own] - bg[s.cells[:n_own]]).max():.3e}  "
      f"diag err {np.abs(diag - dgl[s.cells[:n_own]]).max():.3e}", flush=True)
if bad.size:
    cells = s.cells[:n_own][bad[:10]]
    print(f"[rank {rank}] first bad global cells {cells.tolist()} centers x {g.cell_centers[0, cells].round(3).tolist()}", flush=True)
x, info = kr.solve_local(loc, b_loc[:n_own], diag_own=diag, tol=1e-8, maxiter=1000)
print(f"[rank {rank}] solve {info}", flush=True)
dist.barrier()
dist.destroy_process_group()
