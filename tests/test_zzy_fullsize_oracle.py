"""Parity at the BENCHMARKED sizes against the oracle (round-1 verdict, weak point 2): the reference cannot discretize
998,250 tetrahedra or 128^3 hexahedra here in reasonable time / memory, but a face row of MPFA / MPSA depends only on the
cells around the face's nodes.  So: discretize the full-size mesh on the device, cut seeded patches of a few hundred cells
out of it (one of them on the domain boundary), run the oracle (oracle/fv_oracle.py, pinned to the reference by
tests/test_oracle_vs_golden.py) on each patch, and compare every row whose face has no node on the artificial patch
boundary -- the reference's own split == unsplit argument (applications/test_utils/common_xpfa_tests.py:832-957), with
the oracle on one side.  Rows leave the device through a selection SpGEMM (``S @ M``), not by downloading 24 GB.

CPU leg: the same machinery on small meshes with the host build of the kernels."""
import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200 import shard as sh

TOL = 1e-10
FLOW_KEYS = ("flux", "bound_flux", "bound_pressure_cell", "bound_pressure_face", "vector_source")
MECH_KEYS = ("stress", "bound_stress", "bound_displacement_cell", "bound_displacement_face")


def _params(g, seed=3):
    rng = np.random.default_rng(seed)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bf = g.get_all_boundary_faces()
    x = g.face_centers[:, bf]
    lab = np.where(x[0] < 1e-10, "dir", np.where(x[1] < 1e-10, "rob", "neu"))
    bc = pb.BoundaryCondition(g, bf, list(lab))
    bc.robin_weight = 0.5 + rng.random(g.num_faces)
    C = pb.FourthOrderTensor(np.exp(0.5 * rng.standard_normal(nc)), np.exp(0.5 * rng.standard_normal(nc)))
    vbc = pb.BoundaryConditionVectorial(g, bf[x[0] < 1e-10], "dir")
    return k, bc, C, vbc


def _patch_cells(g, lo, hi):
    """Cells whose centre lies in the box [lo, hi)."""
    cc = g.cell_centers
    return np.flatnonzero(np.all((cc >= np.asarray(lo)[:, None]) & (cc < np.asarray(hi)[:, None]), axis=0))


def _comparable_faces(shard):
    """Faces of the patch none of whose nodes lies on an artificial (cut) face."""
    lg = shard.grid
    fn = sps.csc_matrix(lg.face_nodes)
    bad_node = np.zeros(lg.num_nodes, bool)
    cut = np.flatnonzero(shard.cut_face)
    for f in cut:
        bad_node[fn.indices[fn.indptr[f]:fn.indptr[f + 1]]] = True
    touches = np.add.reduceat(bad_node[fn.indices].astype(np.int64), fn.indptr[:-1]) > 0
    return ~touches


def _rows(M, rows, on_device):
    """Rows ``rows`` of the discretization matrix ``M`` as scipy CSR."""
    if not on_device:
        return sps.csr_matrix(M)[rows]
    from porepy_b200 import ad
    from porepy_b200.sparse import DeviceCsr
    S = sps.csr_matrix((np.ones(rows.size), (np.arange(rows.size), rows)), shape=(rows.size, M.shape[0]))
    return DeviceCsr(S).matmul(ad.as_device_csr(M)).to_scipy()


def check_patches(g, boxes, on_device, seed=3):
    from oracle import fv_oracle as fo
    k, bc, C, vbc = _params(g, seed)
    flow = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    mech = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc})
    pb.Mpfa("flow").discretize(g, flow)
    pb.Mpsa("mech").discretize(g, mech)
    got = {**flow[pb.DISCRETIZATION_MATRICES]["flow"], **mech[pb.DISCRETIZATION_MATRICES]["mech"]}
    eta = pb.determine_eta(g)
    nrows_checked, worst = 0, (0.0, None)
    for lo, hi in boxes:
        cells = _patch_cells(g, lo, hi)
        assert cells.size > 0
        shard = sh.extract_cells(g, cells, np.ones(g.num_faces, bool), np.ones(g.num_cells, bool))
        ok = _comparable_faces(shard)
        assert ok.any(), "patch too small: every face touches its artificial boundary"
        shard.own_face = ok
        lg = shard.grid
        ref = dict(fo.mpfa(lg, k.values[:, :, cells], sh.restrict_scalar_bc(bc, shard), eta))
        ref.update(fo.mpsa(lg, C.values[:, :, cells], sh.restrict_vector_bc(vbc, shard), eta))
        faces = shard.faces[ok]
        for key in FLOW_KEYS + MECH_KEYS:
            br = 1 if key in FLOW_KEYS else 3
            rows = (faces[:, None] * br + np.arange(br)[None, :]).ravel()
            want = sh.embed(shard, key, ref[key])[rows]
            have = _rows(got[key], rows, on_device)
            scale = abs(want).max() if want.nnz else 1.0
            diff = abs(want - have)
            err = float(diff.max() / scale) if diff.nnz else 0.0
            worst = max(worst, (err, key))
            assert err < TOL, (key, err, lo, hi)
            nrows_checked += rows.size
    return nrows_checked, worst


@pytest.fixture()
def host_build(monkeypatch):
    from emu_binding import EmuBackedPlan
    from porepy_b200 import fv
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)


@pytest.mark.parametrize("make,boxes", [
    (lambda: pb.structured_tet_grid([6, 6, 6]), [((0.0, 0.0, 0.0), (0.5, 0.5, 0.5)), ((1 / 6, 1 / 3, 1 / 6), (5 / 6, 1.0, 5 / 6))]),
    (lambda: pb.cart_grid_3d([8, 8, 8]), [((0.0, 0.0, 0.25), (0.5, 0.5, 0.75)), ((0.25, 0.25, 0.25), (0.875, 0.875, 0.875))]),
])
def test_patch_rows_match_the_oracle_host_build(make, boxes, host_build):
    n, worst = check_patches(make(), boxes, on_device=False)
    assert n > 200, n


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dims,boxes", [
    ("tet", (55, 55, 55), [((0.0, 0.0, 0.0), (4 / 55, 4 / 55, 4 / 55)),                 # a corner: Dirichlet + Robin faces
                           ((20 / 55, 30 / 55, 10 / 55), (24 / 55, 34 / 55, 14 / 55)),
                           ((50 / 55, 25 / 55, 51 / 55), (54 / 55, 29 / 55, 1.0))]),
    ("cart", (128, 128, 128), [((0.0, 0.0, 0.0), (5 / 128, 5 / 128, 5 / 128)),
                               ((60 / 128, 70 / 128, 40 / 128), (66 / 128, 76 / 128, 46 / 128)),
                               ((120 / 128, 64 / 128, 122 / 128), (126 / 128, 70 / 128, 1.0))]),
])
def test_benchmark_size_rows_match_the_oracle(kind, dims, boxes):
    g = pb.structured_tet_grid(dims) if kind == "tet" else pb.cart_grid_3d(dims)
    n, worst = check_patches(g, boxes, on_device=True)
    assert n > 1000, n
    print(f"{kind} {dims}: {n} rows of 9 matrices against the oracle, worst relative error {worst}")
