"""``porepy_b200.mdgrid``: the synthetic mixed-dimensional meshes of the bench (3-D hexahedra / structured tetrahedra cut
by disjoint planar fractures).  The hexahedral case is pinned to the reference's mesher in tests/test_porepy_plugin.py
(build container); here: topology invariants on both cell types and exactness of the coupled problem on them."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200 import mdgrid
from porepy_b200.mdflow import MdInterface, MdSubdomain, MixedDimensionalFlow


def network(kind, n=4):
    g = pb.cart_grid_3d([n, n, n]) if kind == "hex" else pb.structured_tet_grid([n, n, n])
    h = 1.0 / n
    sets = [mdgrid.faces_on_rectangle(g, 0, h, (h, h), (1 - h, 1 - h)),
            mdgrid.faces_on_rectangle(g, 0, 1 - h, (0, 0), (1, 1))]     # the second one reaches the domain boundary
    return g, sets, mdgrid.split_fractures(g, sets)


def problem(net, a=1e-2, kn=5.0, value=lambda x: 2.0 + 0 * x[0]):
    m = net.matrix
    bf = m.get_boundary_faces()
    bcv = np.zeros(m.num_faces)
    bcv[bf] = value(m.face_centers[:, bf])
    rng = np.random.default_rng(3)
    k3 = pb.SecondOrderTensor(1 + rng.random(m.num_cells))
    subs = [MdSubdomain(m, pb.initialize_data({}, "flow", {"second_order_tensor": k3,
                                                           "bc": pb.BoundaryCondition(m, bf, "dir")}), bcv)]
    intfs = []
    for k, fg in enumerate(net.fractures):
        db = fg.get_boundary_faces()
        fb = np.zeros(fg.num_faces)
        fb[db] = value(fg.face_centers[:, db])
        subs.append(MdSubdomain(fg, pb.initialize_data({}, "flow", {
            "second_order_tensor": pb.SecondOrderTensor(a * 100.0 * np.ones(fg.num_cells)),
            "bc": pb.BoundaryCondition(fg, db, "dir"), "ambient_dimension": 3}), fb))
        it = net.interfaces[k]
        intfs.append(MdInterface(0, k + 1, it["mortar_to_primary_int"], it["primary_to_mortar_avg"],
                                 it["mortar_to_secondary_int"], it["secondary_to_mortar_avg"],
                                 np.full(it["cell_volumes"].size, kn), it["cell_volumes"], np.full(fg.num_cells, a)))
    return MixedDimensionalFlow(subs, intfs)


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_split_topology(kind):
    g, sets, net = network(kind)
    m = net.matrix
    nff = sum(s.size for s in sets)
    assert m.num_faces == g.num_faces + nff and m.num_cells == g.num_cells
    assert int(m.tags["fracture_faces"].sum()) == 2 * nff
    ncell = np.asarray(abs(m.cell_faces).sum(axis=1)).ravel()
    assert np.all(ncell[m.tags["fracture_faces"]] == 1) and np.all(ncell[~m.tags["fracture_faces"]
                                                                        & ~m.tags["domain_boundary_faces"]] == 2)
    assert np.all(np.asarray(m.cell_faces.sum(axis=0)).ravel() == 0) or kind == "tet"
    # nodes: the first fracture duplicates its (n-3)^2 interior nodes, the one cutting the whole domain all (n+1)^2
    n = round(g.num_nodes ** (1 / 3)) - 1
    assert m.num_nodes == g.num_nodes + (n - 3) ** 2 + (n + 1) ** 2
    # the two copies of a fracture face: same centre, different node sets wherever a node was split
    for (s0, s1), F in zip(net.sides, sets):
        assert np.array_equal(s0, F) and np.allclose(m.face_centers[:, s0], m.face_centers[:, s1])
    # every cell still has its nodes where they were: face-node coordinates agree with the unsplit grid
    fn_old = g.nodes[:, g.face_nodes.indices]
    fn_new = m.nodes[:, m.face_nodes.indices[:g.face_nodes.indices.size]]
    assert np.array_equal(fn_old, fn_new)
    for fg, F in zip(net.fractures, sets):
        assert fg.num_cells == F.size and np.allclose(fg.cell_volumes, g.face_areas[F])
        per_cell = 4 if kind == "hex" else 3
        assert fg.cell_faces.nnz == per_cell * fg.num_cells
        # divergence-free constant field: sum over the edges of a cell of the outward normals vanishes
        out = fg.face_normals @ fg.cell_faces
        assert np.abs(out).max() < 1e-12
    assert int(net.fractures[0].tags["domain_boundary_faces"].sum()) == 0
    assert int(net.fractures[1].tags["tip_faces"].sum()) == 0


def test_touching_fractures_are_refused():
    g = pb.cart_grid_3d([4, 4, 4])
    a = mdgrid.faces_on_rectangle(g, 0, 0.5, (0, 0), (1, 1))
    b = mdgrid.faces_on_rectangle(g, 1, 0.5, (0, 0), (1, 1))
    with pytest.raises(ValueError, match="share nodes"):
        mdgrid.split_fractures(g, [a, b])
    with pytest.raises(ValueError, match="share faces"):
        mdgrid.split_fractures(g, [a, a])
    with pytest.raises(ValueError, match="interior"):
        mdgrid.split_fractures(g, [np.flatnonzero(g.tags["domain_boundary_faces"])[:3]])


@pytest.fixture()
def host_build(monkeypatch):
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan, emu_interface_upwind_masks
    from porepy_b200 import fv
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    monkeypatch.setattr(fv, "interface_upwind_masks", emu_interface_upwind_masks)


def _solve_and_check(prob, value, tol):
    prob.discretize()
    J, b = prob.assemble_host()
    x = spla.spsolve(J.tocsc(), b)
    ps, lam = prob.split(x)
    for p, s in zip(ps, prob.subdomains):
        assert np.abs(p - value(s.sd.cell_centers)).max() < tol
    return lam


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_constant_pressure_is_reproduced(kind, host_build):
    lam = _solve_and_check(problem(network(kind)[2]), lambda x: 2.0 + 0 * x[0], 1e-10)
    assert max(np.abs(v).max() for v in lam) < 1e-10


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_pressure_linear_along_the_fractures_is_reproduced(kind, host_build):
    """p = 1 + y - 2 z is tangential to the fractures (planes x = const): no jump, no interface flux; with constant
    permeabilities MPFA is exact for it on every subdomain."""
    net = network(kind)[2]
    prob = problem(net, value=lambda x: 1.0 + x[1] - 2.0 * x[2])
    m = net.matrix
    prob.subdomains[0].data[pb.PARAMETERS]["flow"]["second_order_tensor"] = pb.SecondOrderTensor(np.ones(m.num_cells))
    for s in prob.subdomains[1:]:                   # tips: the flux of the exact solution, -k grad p . n_out
        fg = s.sd
        tips = np.flatnonzero(fg.tags["tip_faces"])
        k = s.data[pb.PARAMETERS]["flow"]["second_order_tensor"].values[0, 0, 0]
        out = np.asarray(fg.cell_faces[tips].sum(axis=1)).ravel()
        s.bc_values[tips] = -k * out * (fg.face_normals[1, tips] - 2.0 * fg.face_normals[2, tips])
    lam = _solve_and_check(prob, lambda x: 1.0 + x[1] - 2.0 * x[2], 1e-9)
    assert max(np.abs(v).max() for v in lam) < 1e-9


@pytest.mark.parametrize("kind", ["tet", "cart"])
def test_bench_md_network_block_host_build(kind, host_build, monkeypatch):
    """The mixed-dimensional extra of bench.py on a small lattice (host build, scipy stand-in for the device algebra;
    BiCGStab on the Schur complement through the torch recurrence)."""
    import os
    import sys
    import torch
    import emu_sparse
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    emu_sparse.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{**k, "device": "cpu"}))
    out = bench.md_network_block(kind, (10, 10, 10), solve=True)
    assert out["solve"]["converged"] and out["solve"]["true_relres_full_system"] < 1e-7, out["solve"]
    d = out["problem"]
    assert d["fractures"] >= 8 and d["mortar_cells"] == 2 * d["fracture_cells"]
    assert out["jacobian_rows"] == d["dofs"] == d["matrix_cells"] + d["fracture_cells"] + d["mortar_cells"]
    assert len(out["calls"]) == 3 and out["cells_per_s"] > 0
    nw = out["newton"]
    assert "error" not in nw, nw
    assert nw["history"][-1]["residual"] <= 1e-6 * nw["history"][0]["residual"], nw
    assert all(h.get("linear_converged", True) for h in nw["history"]), nw
    prob, _ = bench.md_network_problem(kind, (10, 10, 10))
    prob.discretize()
    J, b = prob.assemble_host()
    x = spla.spsolve(J.tocsc(), b)
    ps, lam = prob.split(x)
    assert 0.0 <= ps[0].min() and ps[0].max() <= 1.0 + 1e-12          # discrete maximum principle (K diagonal)
    assert sum(np.abs(v).sum() for v in lam) > 0
