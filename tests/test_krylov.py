"""Distributed SpMV + BiCGStab (porepy_b200/krylov.py).  CPU: host logic (row partition, halo
exchange plan, solver recurrences) with a scipy matvec stand-in, single process and 2-rank gloo.
GPU: the device path (DeviceCsr through raw torch device pointers) against scipy's direct solve."""
import os

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200 import krylov as kr
from porepy_b200 import shard as sh


def _flow_system(dims=(7, 6, 5)):
    """A = div @ flux, b from Dirichlet data -- assembled with the oracle (CPU checker)."""
    from oracle import fv_oracle as fo
    g = pb.cart_grid_3d(dims, perturb=0.3, seed=4)
    rng = np.random.default_rng(2)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    bc = pb.BoundaryCondition(g, bf[(x < 1e-10) | (x > 1 - 1e-10)], "dir")
    m = fo.mpfa(g, k.values, bc, 0.0)
    div = g.divergence(1)
    bv = np.zeros(g.num_faces)
    bv[bf[x < 1e-10]] = 1.0
    return g, (div @ m["flux"]).tocsr(), -div @ (m["bound_flux"] @ bv)


def _scipy_matvec(loc):
    import torch
    a = loc.a_local
    return lambda xb: torch.as_tensor(a @ xb.cpu().numpy())


def test_single_process_host_logic():
    g, A, b = _flow_system()
    x, owned, info = kr.solve(A, b, tol=1e-11, device="cpu", matvec_factory=_scipy_matvec)
    ref = spla.spsolve(sps.csc_matrix(A), b)
    assert info["converged"]
    assert np.linalg.norm(x.numpy() - ref[owned]) <= 1e-8 * np.linalg.norm(ref)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, A, b = _flow_system()
    owner = sh.partition_cells(g, world)
    x, owned, info = kr.solve(A, b, owner=owner, tol=1e-11, device="cpu", matvec_factory=_scipy_matvec)
    out = [None] * world if rank == 0 else None
    dist.gather_object((owned, x.numpy(), info), out, dst=0)
    if rank == 0:
        full = np.zeros(A.shape[0])
        for o, xv, _ in out:
            full[o] = xv
        ref = spla.spsolve(sps.csc_matrix(A), b)
        q.put((float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), out[0][2]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_halo_exchange_and_solve():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, info = q.get(timeout=300)
    for p in procs:
        p.join(60)
    assert info["converged"] and info["halo_bytes_per_spmv"] > 0
    assert err < 1e-8


def _shard_flow_rows(g, k, bc, bv, part, rank):
    """Own rows of A = div @ flux and of b, assembled from THIS RANK'S SHARD ONLY (oracle as the CPU stand-in of
    the kernels): what ``DevicePlan.mpfa_system().truncate_rows(n_own)`` / ``mpfa_rhs`` give on the device."""
    from oracle import fv_oracle as fo
    s = sh.extract_shard(g, part, rank)
    kl = s.restrict_cell_array(k.values)
    m = fo.mpfa(s.grid, kl, sh.restrict_scalar_bc(bc, s), 0.0)
    div = s.grid.divergence(1)
    n_own = int(s.own_cell.sum())
    a_rows = (div @ m["flux"]).tocsr()[:n_own]
    b_own = (-div @ (m["bound_flux"] @ bv[s.faces]))[:n_own]
    return s, a_rows, b_own


def _problem(dims=(7, 6, 5)):
    g = pb.cart_grid_3d(dims, perturb=0.3, seed=4)
    rng = np.random.default_rng(2)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    bc = pb.BoundaryCondition(g, bf[(x < 1e-10) | (x > 1 - 1e-10)], "dir")
    bv = np.zeros(g.num_faces)
    bv[bf[x < 1e-10]] = 1.0
    return g, k, bc, bv


def test_shard_rows_are_the_global_rows():
    """The system rows a shard assembles on its own (own cells first, columns [own | ghost]) are the rows of the
    global A = div @ flux, and its own right-hand-side entries those of the global b -- no global matrix needed."""
    g, k, bc, bv = _problem()
    _, A, b = _flow_system()
    part = sh.partition_cells(g, 3)
    for r in range(3):
        s, a_rows, b_own = _shard_flow_rows(g, k, bc, bv, part, r)
        n_own = int(s.own_cell.sum())
        ref = A[s.cells[:n_own]][:, s.cells]
        assert abs(ref - a_rows).max() <= 1e-12 * abs(A).max()
        assert np.allclose(b_own, b[s.cells[:n_own]], atol=1e-13)
        # nothing outside the shard is referenced by the own rows
        outside = np.setdiff1d(np.arange(g.num_cells), s.cells)
        assert abs(A[s.cells[:n_own]][:, outside]).sum() == 0


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, k, bc, bv = _problem()
    part = sh.partition_cells(g, world)
    s, a_rows, b_own = _shard_flow_rows(g, k, bc, bv, part, rank)
    loc = kr.local_system_from_shard(s, part, a_rows)
    x, info = kr.solve_local(loc, b_own, diag_own=a_rows.diagonal(), tol=1e-11, device="cpu",
                             matvec_factory=_scipy_matvec)
    out = [None] * world if rank == 0 else None
    dist.gather_object((loc.owned, x.numpy(), info), out, dst=0)
    if rank == 0:
        _, A, b = _flow_system()
        full = np.zeros(A.shape[0])
        for o, xv, _ in out:
            full[o] = xv
        ref = spla.spsolve(sps.csc_matrix(A), b)
        q.put((float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), out[0][2]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_assembly_and_solve():
    """Each rank: its shard only -> own rows of A and b -> distributed BiCGStab (halo exchange in the shard's
    ghost order).  Solution = direct solve of the unsplit system."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, info = q.get(timeout=300)
    for p in procs:
        p.join(60)
    assert info["converged"] and info["halo_bytes_per_spmv"] > 0
    assert err < 1e-8


def test_local_system_partition_is_consistent():
    g, A, b = _flow_system((5, 4, 3))
    owner = sh.partition_cells(g, 3)
    x = np.random.default_rng(0).standard_normal(A.shape[1])
    y = A @ x
    for r in range(3):
        # world=1 avoids the collective; rebuild the send lists by hand is not needed for this check
        owned = np.flatnonzero(owner == r)
        rows = A[owned]
        loc = kr.build_local_system(A, np.where(owner == r, 0, 1), 0, 1)
        assert loc.a_local.shape == (owned.size, owned.size + loc.ghosts.size)
        xb = np.r_[x[loc.owned], x[loc.ghosts]]
        assert np.allclose(loc.a_local @ xb, y[loc.owned])
        del rows


@pytest.mark.gpu
def test_gpu_bicgstab_matches_direct_solve():
    g = pb.cart_grid_3d([16, 14, 12], perturb=0.3, seed=4)
    rng = np.random.default_rng(2)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    xx = g.face_centers[0, bf]
    bc = pb.BoundaryCondition(g, bf[(xx < 1e-10) | (xx > 1 - 1e-10)], "dir")
    bv = np.zeros(g.num_faces)
    bv[bf[xx < 1e-10]] = 1.0
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc, "bc_values": bv})
    d = pb.Mpfa("flow")
    d.discretize(g, data)
    A, b = d.assemble_matrix_rhs(g, data)
    x, owned, info = kr.solve(A, b, tol=1e-11)
    ref = spla.spsolve(sps.csc_matrix(A), b)
    assert info["converged"], info
    assert np.linalg.norm(x.cpu().numpy() - ref) <= 1e-8 * np.linalg.norm(ref)


def _block_system(nb=60, bs=3, seed=5):
    """Block-structured test matrix: strongly coupled bs x bs diagonal blocks (point Jacobi is a poor preconditioner),
    weak coupling between the blocks."""
    rng = np.random.default_rng(seed)
    blocks = []
    for _ in range(nb):
        q, _r = np.linalg.qr(rng.standard_normal((bs, bs)))
        blocks.append(q @ np.diag(10.0 ** rng.uniform(0, 3, bs)) @ q.T)
    a = sps.block_diag(blocks).tolil()
    n = nb * bs
    for _ in range(4 * n):
        i, j = rng.integers(0, n, 2)
        if i // bs != j // bs:
            a[i, j] += 0.05 * rng.standard_normal()
    return sps.csr_matrix(a), rng.standard_normal(n)


def test_block_jacobi_host_logic():
    """Eager recurrence with the block-Jacobi preconditioner (inverted 3 x 3 diagonal blocks) on a matvec stand-in."""
    import torch
    a, b = _block_system()
    loc = kr.build_local_system(a, np.zeros(a.shape[0], dtype=np.int64), 0, 1)
    op = kr.DistributedOperator(loc, "cpu", matvec=_scipy_matvec(loc))
    blk = np.stack([np.linalg.inv(a[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray()) for i in range(a.shape[0] // 3)])
    x, info = kr.bicgstab(op, torch.as_tensor(b), tol=1e-11, maxiter=500, block_inv=(torch.as_tensor(blk.ravel()), 3))
    ref = spla.spsolve(sps.csc_matrix(a), b)
    assert info["converged"], info
    assert np.linalg.norm(x.numpy() - ref) <= 1e-8 * np.linalg.norm(ref)
    _, info_pt = kr.bicgstab(op, torch.as_tensor(b), tol=1e-11, maxiter=500, diag_own=torch.as_tensor(a.diagonal()))
    assert info["iterations"] <= info_pt["iterations"]


@pytest.mark.gpu
def test_gpu_block_jacobi_fused_mechanics_solve():
    """The mechanics system A = div_nd @ stress assembled on the device, solved by the fused BiCGStab with the
    block-Jacobi preconditioner (``pb_csr_block_diag_inv_dev``: one inverted 3 x 3 block per cell), against scipy's
    direct solve; the device block inverses against NumPy."""
    import torch
    g = pb.structured_tet_grid([7, 6, 5])
    rng = np.random.default_rng(3)
    nc = g.num_cells
    C = pb.FourthOrderTensor(np.exp(0.5 * rng.standard_normal(nc)), np.exp(0.5 * rng.standard_normal(nc)))
    bf = g.get_all_boundary_faces()
    vbc = pb.BoundaryConditionVectorial(g, bf[g.face_centers[2, bf] < 1e-10], "dir")
    src = rng.standard_normal(3 * nc) * np.repeat(g.cell_volumes, 3)
    data = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc, "bc_values": np.zeros(3 * g.num_faces),
                                           "source": src})
    d = pb.Mpsa("mech")
    d.discretize(g, data)
    A, b = d.assemble_matrix_rhs(g, data)
    a_dev = A.device_csr
    blk = a_dev.block_diagonal_inverse(3)
    As = sps.csr_matrix(A)
    want = np.stack([np.linalg.inv(As[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray()) for i in range(nc)])
    assert np.abs(blk.cpu().numpy().reshape(-1, 3, 3) - want).max() <= 1e-12 * np.abs(want).max()
    loc = kr.LocalSystem(0, 1, np.arange(3 * nc), np.zeros(0, np.int64), a_dev, [0], [np.zeros(0, np.int64)])
    x, info = kr.solve_local(loc, b, tol=1e-11, maxiter=4000, block_inv=(blk, 3))
    ref = spla.spsolve(sps.csc_matrix(As), b)
    assert info["converged"] and info.get("fused"), info
    assert np.linalg.norm(x.cpu().numpy() - ref) <= 1e-7 * np.linalg.norm(ref)


_M3 = np.array([[2.0, 0.3, 0.0], [0.3, 2.0, 0.1], [0.0, 0.1, 2.0]])


def _shard_worker_dof3(rank, world, port, q):
    """Three unknowns per cell (the layout of the mechanics system): A (x) M3 from the rank's shard rows, block-Jacobi
    preconditioner from the own diagonal blocks, the cell-level halo plan expanded by ``dof=3``."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, k, bc, bv = _problem()
    part = sh.partition_cells(g, world)
    s, a_rows, b_own = _shard_flow_rows(g, k, bc, bv, part, rank)
    a3 = sps.kron(a_rows, _M3).tocsr()
    b3 = np.kron(b_own, np.array([1.0, 2.0, 3.0]))
    loc = kr.local_system_from_shard(s, part, a3, dof=3)
    n_own = b_own.size
    blk = np.stack([np.linalg.inv(a_rows[i, i] * _M3) for i in range(n_own)])
    x, info = kr.solve_local(loc, b3, tol=1e-11, device="cpu", matvec_factory=_scipy_matvec,
                             block_inv=(torch.as_tensor(blk.ravel()), 3))
    out = [None] * world if rank == 0 else None
    dist.gather_object((loc.owned, x.numpy(), info), out, dst=0)
    if rank == 0:
        _, A, b = _flow_system()
        full = np.zeros(3 * A.shape[0])
        for o, xv, _ in out:
            full[o] = xv
        ref = spla.spsolve(sps.csc_matrix(sps.kron(A, _M3)), np.kron(b, np.array([1.0, 2.0, 3.0])))
        q.put((float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), out[0][2]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_three_dof_per_cell_block_jacobi():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker_dof3, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, info = q.get(timeout=300)
    for p in procs:
        p.join(60)
    assert info["converged"] and info["halo_bytes_per_spmv"] > 0
    assert err < 1e-8
