"""The reference's tutorial ``tutorials/stress_discretization.ipynb`` (cells 3-18) with this package's
classes: 5 x 5 Cartesian grid of side 5, unit Lame parameters, clamped bottom, unit traction on the
top; the tutorial asserts that the discrete tractions on the Neumann faces reproduce the data.  On CPU
the device plan is the host build of the kernels."""
import numpy as np

import porepy_b200 as pb
from porepy_b200 import fv


def test_tutorial_traction_identity(monkeypatch):
    from emu_binding import EmuBackedPlan
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    n = 5
    g = pb.cart_grid_2d([n, n], [n, n])
    C = pb.FourthOrderTensor(np.ones(g.num_cells), np.ones(g.num_cells))
    dirich = np.ravel(np.argwhere(g.face_centers[1] < 1e-10))
    bound = pb.BoundaryConditionVectorial(g, dirich, ["dir"] * dirich.size)
    top_faces = np.ravel(np.argwhere(g.face_centers[1] > n - 1e-10))
    u_b = np.zeros((g.dim, g.num_faces))
    u_b[1, top_faces] = -1 * g.face_areas[top_faces]
    u_b = u_b.ravel("F")
    keyword = "mechanics"
    mpsa_class = pb.Mpsa(keyword)
    data = pb.initialize_data({}, keyword, {"fourth_order_tensor": C, "source": np.zeros(g.dim * g.num_cells),
                                            "bc": bound, "bc_values": u_b})
    mpsa_class.discretize(g, data)
    A, b = mpsa_class.assemble_matrix_rhs(g, data)
    u = np.linalg.solve(A.toarray(), b)
    M = data[pb.DISCRETIZATION_MATRICES][keyword]
    T = M[mpsa_class.stress_matrix_key] @ u + M[mpsa_class.bound_stress_matrix_key] @ u_b
    T2d = np.reshape(T, (g.dim, -1), order="F")
    u_b2d = np.reshape(u_b, (g.dim, -1), order="F")
    is_neu = np.asarray(bound.is_neu)[: g.dim]
    assert np.allclose(np.abs(u_b2d[is_neu]), np.abs(T2d[is_neu]))
    assert u[1::2].min() < 0 and abs(u[1::2].min()) > abs(u[0::2]).max()   # compressed downwards
