"""Host-side sharding logic (porepy_b200/shard.py): split == unsplit, the property the
reference pins in applications/test_utils/common_xpfa_tests.py:832-957.  On CPU the local
discretizer is the oracle (the checker standing in for the kernel); the gloo test runs the
N > 1 path with world_size 2.  The GPU version of the same property is in
tests/test_gpu_parity.py::test_sharded_equals_unsplit."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from cases import flatten
from golden_io import rel_err
from oracle import fv_oracle as fo
from porepy_b200 import shard as sh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MPFA_SHAPES = {"flux": ("face", "cell", 1, 1), "bound_flux": ("face", "face", 1, 1),
               "bound_pressure_cell": ("face", "cell", 1, 1), "bound_pressure_face": ("face", "face", 1, 1),
               "vector_source": ("face", "cell", 1, 3), "bound_pressure_vector_source": ("face", "cell", 1, 3)}
MPSA_SHAPES = {"stress": ("face", "cell", 3, 3), "bound_stress": ("face", "face", 3, 3),
               "bound_displacement_cell": ("face", "cell", 3, 3), "bound_displacement_face": ("face", "face", 3, 3),
               "displacement_divergence:a": ("cell", "cell", 1, 3),
               "boundary_displacement_divergence:a": ("cell", "face", 1, 3),
               "scalar_gradient:a": ("face", "cell", 3, 1), "mpsa_consistency:a": ("cell", "cell", 1, 1),
               "bound_displacement_pressure:a": ("face", "cell", 3, 1)}


def _problem(kind):
    g = pb.structured_tet_grid([3, 2, 2]) if kind == "tet" else pb.cart_grid_3d([6, 3, 3], perturb=0.3)
    rng = np.random.default_rng(3)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    bc = pb.BoundaryCondition(g, bf[(x < 1e-10) | (x > 1 - 1e-10)], "dir")
    C = pb.FourthOrderTensor(np.exp(0.3 * rng.standard_normal(nc)), np.exp(0.3 * rng.standard_normal(nc)))
    vbc = pb.BoundaryConditionVectorial(g, bf[g.face_centers[2, bf] < 1e-10], "dir")
    al = 0.5 + rng.random(nc)
    return g, k, bc, C, vbc, al


def shard_discretize(g, k, bc, C, vbc, al, part, rank):
    s = sh.extract_shard(g, part, rank)
    eta = pb.determine_eta(g)
    lf = fo.mpfa(s.grid, s.restrict_cell_array(k.values), sh.restrict_scalar_bc(bc, s), eta)
    alpha = np.eye(3)[:, :, None] * s.restrict_cell_array(al)
    lm = flatten(fo.mpsa(s.grid, s.restrict_cell_array(C.values), sh.restrict_vector_bc(vbc, s), eta,
                         alpha={"a": alpha}))
    out = {key: s.to_global(lf[key], *MPFA_SHAPES[key]) for key in MPFA_SHAPES}
    out.update({key: s.to_global(lm[key], *MPSA_SHAPES[key]) for key in MPSA_SHAPES})
    return out


@pytest.mark.parametrize("kind", ["cart", "tet"])
@pytest.mark.parametrize("nparts", [2, 3])
def test_split_equals_unsplit(kind, nparts):
    g, k, bc, C, vbc, al = _problem(kind)
    eta = pb.determine_eta(g)
    ref = dict(fo.mpfa(g, k.values, bc, eta))
    ref.update(flatten(fo.mpsa(g, C.values, vbc, eta, alpha={"a": np.eye(3)[:, :, None] * al})))
    part = sh.partition_cells(g, nparts)
    acc = None
    owned_faces = np.zeros(g.num_faces, int)
    for r in range(nparts):
        out = shard_discretize(g, k, bc, C, vbc, al, part, r)
        s = sh.extract_shard(g, part, r)
        owned_faces[s.faces[s.own_face]] += 1
        acc = out if acc is None else {key: acc[key] + out[key] for key in out}
    assert np.all(owned_faces == 1)  # every face row produced exactly once
    for key in ref:
        assert rel_err(ref[key], acc[key]) < 1e-12, key


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, k, bc, C, vbc, al = _problem("cart")
    part = sh.partition_cells(g, world)
    out = shard_discretize(g, k, bc, C, vbc, al, part, rank)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object({key: (m.data, m.indices, m.indptr, m.shape) for key, m in out.items()},
                       gathered, dst=0)
    if rank == 0:
        acc = None
        for o in gathered:
            mats = {key: sps.csr_matrix((d, i, p), shape=s) for key, (d, i, p, s) in o.items()}
            acc = mats if acc is None else {key: acc[key] + mats[key] for key in mats}
        eta = pb.determine_eta(g)
        ref = dict(fo.mpfa(g, k.values, bc, eta))
        ref.update(flatten(fo.mpsa(g, C.values, vbc, eta, alpha={"a": np.eye(3)[:, :, None] * al})))
        q.put(max(rel_err(ref[key], acc[key]) for key in ref))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_split_equals_unsplit():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(60)
    assert err < 1e-12


@pytest.mark.parametrize("kind,dims,nparts", [("tet", (5, 4, 3), 3), ("cart", (7, 5, 4), 4), ("cart2d", (9, 7), 2)])
def test_native_extraction_equals_numpy_restatement(kind, dims, nparts):
    """``pb_shard_create`` (csrc/shard.cu, the routine the ranks run) against the NumPy restatement of the same rule,
    field by field: index maps, kept / cut / own flags, the local CSC arrays, the geometry and the tags."""
    if kind == "tet":
        g = pb.structured_tet_grid(list(dims))
    elif kind == "cart":
        g = pb.cart_grid_3d(list(dims), perturb=0.2)
    else:
        g = pb.cart_grid_2d(list(dims))
    part = sh.partition_cells(g, nparts)
    for rank in range(nparts):
        a, b = sh.extract_shard(g, part, rank), sh.extract_shard_numpy(g, part, rank)
        for key in ("cells", "faces", "nodes", "own_cell", "own_face", "cut_face", "own_node"):
            assert np.array_equal(getattr(a, key), getattr(b, key)), key
        assert a.num_global == b.num_global
        for key in ("nodes", "face_normals", "face_centers", "face_areas", "cell_centers", "cell_volumes"):
            assert np.array_equal(getattr(a.grid, key), getattr(b.grid, key)), key
        for key in ("cell_faces", "face_nodes"):
            ma, mb = getattr(a.grid, key), getattr(b.grid, key)
            assert ma.shape == mb.shape and np.array_equal(ma.indptr, mb.indptr)
            assert np.array_equal(ma.indices, mb.indices) and np.array_equal(ma.data, mb.data), key
        for key in b.grid.tags:
            assert np.array_equal(a.grid.tags[key], b.grid.tags[key]), key
        assert (a.grid.dim, a.grid.num_cells, a.grid.num_faces, a.grid.num_nodes) == \
               (b.grid.dim, b.grid.num_cells, b.grid.num_faces, b.grid.num_nodes)
