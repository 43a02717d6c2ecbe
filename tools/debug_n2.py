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
import ctypes as C  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", local)
b_t = torch.as_tensor(b_loc[:n_own], device=dev)
d_t = torch.as_tensor(diag, device=dev)


def gdot(a, b):
    v = torch.dot(a, b).reshape(1).clone()
    dist.all_reduce(v)
    return float(v)


def stepper(iters, sync_mask, verify):
    """The fused iteration launched by hand.  sync_mask bits: 1 = device sync after every vector kernel, 2 = after every
    exchange, 4 = after every all-reduce, 8 = after every SpMV.  verify: every stage is compared with torch arithmetic."""
    n, ng = op.n_own, op.n_ghost
    vec = lambda m=n: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
    x, r, rhat, p, v, s_, t = (vec() for _ in range(7))
    xb_p, xb_s = vec(n + ng), vec(n + ng)
    ph, sh_ = xb_p[:n], xb_s[:n]
    minv = (1.0 / d_t).contiguous()
    scal = torch.zeros(14, dtype=torch.float64, device=dev)
    P = lambda a: C.c_void_p(a.data_ptr())  # noqa: E731
    S = lambda i: C.c_void_p(scal.data_ptr() + 8 * i)  # noqa: E731
    st = torch.cuda.current_stream().cuda_stream
    csr = op.dev_csr
    carry = 1 if rank == 0 else 0
    first_bad = []

    def sync(bit):
        if sync_mask & bit:
            torch.cuda.synchronize()

    def red(lo, hi):
        dist.all_reduce(scal[lo:hi])
        sync(4)

    def chk(tag, it, got, want):
        if not verify or first_bad:
            return
        e = float((got - want).abs().max()) if torch.is_tensor(got) else abs(got - want)
        ref = float(want.abs().max()) if torch.is_tensor(want) else abs(want)
        if e > 1e-9 * max(ref, 1e-300):
            first_bad.append((it, tag, e, ref))
            print(f"\n[rank {rank}] FIRST MISMATCH it {it} stage {tag}: err {e:.3e} ref {ref:.3e}", flush=True)

    _lib.check(lib.pb_kry_init(n, P(b_t), P(x), P(r), P(rhat), P(p), P(v), P(scal), 1e-8, st))
    red(10, 11)
    _lib.check(lib.pb_kry_seed(P(scal), st))
    bb = float(scal[10])
    hist = []
    for it in range(iters):
        cur = it & 1
        g_ = 5 * cur
        gp = 5 * (cur ^ 1)
        if verify:
            h = scal.cpu().numpy().copy()
            alpha_prev, omega_prev = h[gp + 4] / h[gp + 0], h[gp + 1] / h[gp + 2]
            beta = (h[g_ + 4] / h[gp + 4]) * (alpha_prev / omega_prev)
            p_want = r + beta * (p - omega_prev * v)
            chk("rho(cur) == (rhat, r)", it, float(h[g_ + 4]), gdot(rhat, r))
            chk("rr(cur) == (r, r)", it, float(h[g_ + 3]), gdot(r, r))
        _lib.check(lib.pb_kry_p(n, P(r), P(p), P(v), P(minv), P(ph), P(scal), cur, 1, st))
        sync(1)
        if verify:
            chk("p", it, p, p_want)
            chk("ph", it, ph, p_want * minv)
        op.exchange_into(xb_p)
        sync(2)
        if verify:
            want = op.exchange(ph.clone()).clone()
            chk("halo(p)", it, xb_p, want)
        _lib.check(lib.pb_csr_spmv_dots_dev(csr.h, P(xb_p), P(v), P(rhat), S(g_ + 0), None, None, st))
        sync(8)
        if verify:
            v_want = op.matvec(ph.clone())
            chk("v = A ph", it, v, v_want)
            loc_dot = float(torch.dot(rhat, v))
            chk("local (rhat, v)", it, float(scal[g_ + 0]), loc_dot)
        red(g_ + 0, g_ + 1)
        if verify:
            chk("reduced (rhat, v)", it, float(scal[g_ + 0]), gdot(rhat, v))
            h = scal.cpu().numpy().copy()
            alpha = h[g_ + 4] / h[g_ + 0]
            s_want = r - alpha * v
        _lib.check(lib.pb_kry_s(n, P(r), P(v), P(minv), P(s_), P(sh_), P(scal), cur, 1, st))
        sync(1)
        if verify:
            chk("s", it, s_, s_want)
            chk("sh", it, sh_, s_want * minv)
            chk("prev group zeroed", it, float(scal[gp:gp + 5].abs().max()), 0.0)
        op.exchange_into(xb_s)
        sync(2)
        if verify:
            chk("halo(s)", it, xb_s, op.exchange(sh_.clone()).clone())
        _lib.check(lib.pb_csr_spmv_dots_dev(csr.h, P(xb_s), P(t), P(s_), S(g_ + 1), None, S(g_ + 2), st))
        sync(8)
        if verify:
            chk("t = A sh", it, t, op.matvec(sh_.clone()))
        red(g_ + 1, g_ + 3)
        if verify:
            chk("reduced (t, s)", it, float(scal[g_ + 1]), gdot(t, s_))
            chk("reduced (t, t)", it, float(scal[g_ + 2]), gdot(t, t))
            h = scal.cpu().numpy().copy()
            alpha, omega = h[g_ + 4] / h[g_ + 0], h[g_ + 1] / h[g_ + 2]
            x_want = x + alpha * ph + omega * sh_
            r_want = s_ - omega * t
        _lib.check(lib.pb_kry_xr(n, P(x), P(ph), P(sh_), P(s_), P(t), P(r), P(rhat), P(scal), cur, carry, st))
        sync(1)
        if verify:
            chk("x", it, x, x_want)
            chk("r", it, r, r_want)
        red(gp + 3, gp + 5)
        if verify or (it % 20) == 19 or it == iters - 1:
            h = scal.cpu().numpy().copy()
            rel = float(np.sqrt(max(h[gp + 3], 0.0) / bb))
            hist.append((it + 1, rel))
            if verify:
                true_r = b_t - op.matvec((x).clone())
                rel_true = float(np.sqrt(gdot(true_r, true_r) / bb))
                hist[-1] = (it + 1, rel, rel_true)
            if h[11] != 0.0:
                break
    both = [None] * world if rank == 0 else None
    dist.gather_object((scal.cpu().numpy().tolist(), hist[-3:]), both, dst=0)
    if rank == 0:
        same = both[0][0][:11] == both[1][0][:11]
        print(f"\n[rank 0] stepper sync_mask {sync_mask} verify {verify}: scal identical on both ranks {same}; history tail {both[0][1]}"
              f"  first mismatch {first_bad[:1]}", flush=True)
        if not same:
            print("\n[rank 0]   r0", both[0][0], "\n[rank 0]   r1", both[1][0], flush=True)


stepper(30, 15, True)
stepper(400, 0, False)
os.environ["POREB200_KRYLOV_TRACE"] = "1"
for graph in ("0", "1"):
    os.environ["POREB200_KRYLOV_GRAPH"] = graph
    xf, inf = kr.bicgstab(op, b_t, tol=1e-8, maxiter=600, diag_own=d_t)
    tr = inf.pop("trace", [])
    if rank == 0:
        print(f"\n[rank 0] distributed fused graph={graph}: {inf}  tail {[(t[0], float(f'{t[1]:.2e}')) for t in tr[-2:]]}", flush=True)
xe, ine = kr.bicgstab(op, b_t, x0=torch.zeros_like(b_t), tol=1e-8, maxiter=600, diag_own=d_t)
if rank == 0:
    print(f"\n[rank 0] distributed eager: {ine}", flush=True)
    print(f"\n[rank 0] |x_fused - x_eager| / |x| = {float((xf - xe).norm() / xe.norm()):.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
