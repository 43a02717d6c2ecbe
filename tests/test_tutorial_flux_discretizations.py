"""The reference's tutorial ``tutorials/flux_discretizations.ipynb`` (cells 7-30) with this package's
classes: 20 x 20 Cartesian grid of the unit square, unit permeability, unit source, homogeneous
Dirichlet boundary; the tutorial asserts ``sum(p_tpfa) == 14.192684340967551`` and
``sum(p_mpfa) == 14.192684340967542`` (np.isclose).  On CPU the device plan is the host build of the
kernels; tests/test_zz_tutorial_gpu.py runs the MPFA part on the GPU."""
import numpy as np
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200 import fv


def tutorial_problem():
    g = pb.cart_grid_2d([20, 20], [1, 1])
    perm = pb.SecondOrderTensor(np.ones(g.num_cells))
    b_faces = g.tags["domain_boundary_faces"].nonzero()[0]
    bc = pb.BoundaryCondition(g, b_faces, ["dir"] * b_faces.size)
    parameters = {"second_order_tensor": perm, "source": g.cell_volumes, "bc": bc,
                  "bc_values": np.zeros(g.num_faces)}
    return g, pb.initialize_data({}, "flow", parameters)


def solve(discr, g, data):
    discr.discretize(g, data)
    A, b_flow = discr.assemble_matrix_rhs(g, data)
    return spla.spsolve(A.tocsc(), b_flow + data[pb.PARAMETERS]["flow"]["source"])


def test_tutorial_numbers(monkeypatch):
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    g, data = tutorial_problem()
    p_tpfa = solve(pb.Tpfa("flow"), g, data)
    assert np.isclose(np.sum(p_tpfa), 14.192684340967551)
    g, data = tutorial_problem()
    p_mpfa = solve(pb.Mpfa("flow"), g, data)
    assert np.isclose(np.sum(p_mpfa), 14.192684340967542)
    # isotropic K on an orthogonal grid: the two schemes coincide
    assert np.abs(p_tpfa - p_mpfa).max() < 1e-12
