"""CPU checks of the host side: the C-ABI library loads and exports every symbol that
include/poreb200.h declares, pattern expansion, mesh generators, parameter mirrors, and the
loud failure without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200 import _lib, build
from porepy_b200.fv import block_expand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "poreb200.h")).read()
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"libporeb200.so does not export {name}"
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_no_cpu_fallback(lib):
    if lib.pb_device_count() >= 1:
        pytest.skip("GPU present")
    g = pb.cart_grid_3d([2, 2, 2])
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor(np.ones(8)),
                                           "bc": pb.BoundaryCondition(g)})
    with pytest.raises(RuntimeError):
        pb.Mpfa("flow").discretize(g, data)
    with pytest.raises(RuntimeError):
        pb.DeviceCsr(sps.identity(4, format="csr"))


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import porepy_b200, porepy_b200.fv, porepy_b200.sparse; "
            "bad=[m for m in sys.modules if m.startswith('oracle') or 'emu' in m]; "
            "assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_block_expand_layout():
    ip = np.array([0, 2, 3], np.int32)
    ix = np.array([0, 2, 1], np.int32)
    nip, nix = block_expand(ip, ix, 2, 3)
    assert nip.tolist() == [0, 6, 12, 15, 18]
    assert nix[:6].tolist() == [0, 1, 2, 6, 7, 8] and nix[6:12].tolist() == [0, 1, 2, 6, 7, 8]
    assert nix[12:15].tolist() == [3, 4, 5]
    # position rule: br*bc*ip[r] + i*bc*len + (p-ip[r])*bc + j
    r, i, p, j = 0, 1, 1, 2
    pos = 2 * 3 * ip[r] + i * 3 * 2 + (p - ip[r]) * 3 + j
    assert nix[pos] == ix[p] * 3 + j and nip[r * 2 + i] <= pos < nip[r * 2 + i + 1]


@pytest.mark.parametrize("make", [lambda: pb.cart_grid_3d([3, 2, 4], perturb=0.3),
                                  lambda: pb.structured_tet_grid([2, 3, 2])])
def test_mesh_generators_are_consistent(make):
    g = make()
    # closed cells: sum of outward normals vanishes; volumes fill the box
    assert abs(g.cell_faces.T @ g.face_normals.T).max() < 1e-13
    assert np.isclose(g.cell_volumes.sum(), 1.0)
    cf = sps.coo_matrix(g.cell_faces)
    out = np.einsum("ij,ij->j", g.face_normals[:, cf.row] * cf.data,
                    g.face_centers[:, cf.row] - g.cell_centers[:, cf.col])
    assert (out > 0).all()
    assert np.allclose(np.linalg.norm(g.face_normals, axis=0), g.face_areas)
    # divergence theorem for the linear field x: sum_f sgn * n_f . x_f = 3 V
    flux = np.einsum("ij,ij->j", g.face_normals, g.face_centers)
    assert np.allclose(g.cell_faces.T @ flux, 3 * g.cell_volumes)
    nb = g.tags["domain_boundary_faces"].sum()
    assert nb == (np.asarray(abs(g.cell_faces).sum(axis=1)).ravel() == 1).sum()


def test_fourth_order_tensor_layout():
    t = pb.FourthOrderTensor(np.array([2.0]), np.array([3.0]))
    v = t.values[:, :, 0]
    assert v[0, 0] == 2 * 2 + 3 and v[0, 4] == 3 and v[1, 1] == 2 and v[1, 3] == 2 and v[4, 8] == 3
    assert np.allclose(v, v.T)


def test_boundary_condition_defaults():
    g = pb.cart_grid_3d([2, 2, 2])
    bc = pb.BoundaryCondition(g, g.get_all_boundary_faces()[:3], "dir")
    assert bc.is_dir.sum() == 3 and bc.is_neu.sum() == g.get_all_boundary_faces().size - 3
    vb = pb.BoundaryConditionVectorial(g)
    assert vb.is_neu.shape == (3, g.num_faces) and vb.robin_weight.shape == (3, 3, g.num_faces)


def test_bc_encoding_and_error_behaviour():
    """Host-side encoding of the reference's BC objects and its exception parity."""
    from porepy_b200.fv import scalar_bc_codes, vector_bc_codes
    g = pb.cart_grid_3d([2, 2, 2])
    bf = g.get_all_boundary_faces()
    bc = pb.BoundaryCondition(g, bf[:4], ["dir", "rob", "neu", "dir"])
    bc.is_internal[bf[3]] = True  # internal (fracture) faces are Neumann for MPFA (mpfa.py:1452-1454)
    codes = scalar_bc_codes(bc, g.num_faces)
    assert codes[bf[0]] == 1 and codes[bf[1]] == 3 and codes[bf[2]] == 2 and codes[bf[3]] == 2
    assert (codes[np.setdiff1d(np.arange(g.num_faces), bf)] == 0).all()
    vb = pb.BoundaryConditionVectorial(g, bf[:2], ["dir", "rob"])
    vcodes, robw = vector_bc_codes(vb, 3, g.num_faces)
    assert vcodes.shape == (3, g.num_faces) and (vcodes[:, bf[0]] == 1).all() and robw.shape == (3, 3, g.num_faces)
    with pytest.raises(AttributeError):  # mpsa.py:823: "MPSA must be given a vectorial boundary condition"
        vector_bc_codes(bc, 3, g.num_faces)
    from porepy_b200.fv import vector_bc_basis
    assert vector_bc_basis(vb, 3) is None            # identity everywhere
    assert vector_bc_basis(vb, 3, vcodes) is None
    interior = np.setdiff1d(np.arange(g.num_faces), bf)[0]
    vb.basis[0, 1, interior] = 0.5                   # a basis entry on an interior face never enters an equation:
    assert vector_bc_basis(vb, 3, vcodes) is None    # the boundary-only test (what discretize() uses) ignores it
    vb.basis[0, 1, interior] = 0.0
    vb.basis[0, 1, bf[0]] = 0.5
    assert vector_bc_basis(vb, 3).shape == (3, 3, g.num_faces)
    assert vector_bc_basis(vb, 3, vcodes).shape == (3, 3, g.num_faces)


def test_determine_eta_follows_the_reference_rule():
    assert pb.determine_eta(pb.cart_grid_3d([1, 1, 1])) == 0.0
    assert pb.determine_eta(pb.structured_tet_grid([1, 1, 1])) == pytest.approx(1.0 / 3.0)


def test_biot_class_keys_and_assembly_refusal():
    b = pb.Biot("mech")
    keys = {k: v for k, v in vars(b).items() if k.endswith("_matrix_key")}
    assert keys["bound_displacement_divergence_matrix_key"] == "boundary_displacement_divergence"
    assert keys["consistency_matrix_key"] == "mpsa_consistency"
    with pytest.raises(NotImplementedError):  # biot.py:125-149
        b.assemble_matrix_rhs(None, {})


def test_periodic_grids_are_refused():
    """Periodic face pairs are merged by the reference's SubcellTopology (_fvutils.py:95-140);
    the topology plan does not, so the operator must refuse them (before touching the device)."""
    g = pb.cart_grid_3d([2, 2, 2])
    g.periodic_face_map = np.array([[0], [2]])
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor(np.ones(8)),
                                           "bc": pb.BoundaryCondition(g)})
    with pytest.raises(NotImplementedError):
        pb.Mpfa("flow").discretize(g, data)
