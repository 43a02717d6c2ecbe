"""Partial (re)discretization through ``specified_cells / specified_faces / specified_nodes``
(reference _fvutils.py:308-355,1260-1462; mpfa.py:176-201,468-508; biot.py:326-342,614-712), driven on
CPU through the operator classes with the device plan replaced by the host build of the kernels.
The reference's own partial-discretization tests run on the same code through
tools/run_reference_tests.py (tests/test_porepy_plugin.py)."""
import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200 import fv
from cases import flatten
from emu_binding import EmuBackedPlan
from golden_io import rel_err
from test_shard_api import problem


@pytest.fixture(autouse=True)
def emu_plan(monkeypatch):
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)


def _full(g, flow, mech):
    pb.Mpfa("flow").discretize(g, flow)
    pb.Biot("mech").discretize(g, mech)
    out = dict(flow[pb.DISCRETIZATION_MATRICES]["flow"])
    out.update(mech[pb.DISCRETIZATION_MATRICES]["mech"])
    return flatten(out)


def _fresh(data, kw, **extra):
    p = dict(data[pb.PARAMETERS][kw])
    p.update(extra)
    return pb.initialize_data({}, kw, p)


@pytest.mark.parametrize("kind", ["cart", "tet"])
def test_gradual_build_up_by_nodes_equals_the_full_discretization(kind):
    """Nodes mode: every face row is produced by exactly one group of nodes only if all its nodes
    are in the group, so whole-grid coverage needs overlapping groups; the reference's use is to
    sum disjoint FACE sets -- here: each face is taken from the first group that completes it."""
    g, flow, mech = problem(kind)
    ref = _full(g, flow, mech)
    x = g.nodes[0]
    groups = [np.flatnonzero(x < 0.55), np.flatnonzero(x > 0.3)]     # overlapping node slabs
    acc, done = {}, np.zeros(g.num_faces, bool)
    for nodes in groups:
        d1, d2 = _fresh(flow, "flow", specified_nodes=nodes), _fresh(mech, "mech", specified_nodes=nodes)
        pb.Mpfa("flow").discretize(g, d1)
        pb.Biot("mech").discretize(g, d2)
        faces = d1[pb.PARAMETERS]["flow"]["active_faces"]
        assert np.array_equal(faces, d2[pb.PARAMETERS]["mech"]["active_faces"])
        new = np.zeros(g.num_faces, bool)
        new[faces] = True
        new &= ~done
        done |= new
        part = dict(d1[pb.DISCRETIZATION_MATRICES]["flow"])
        part.update(d2[pb.DISCRETIZATION_MATRICES]["mech"])
        for key, m in flatten(part).items():
            if m.shape[0] % g.num_faces or key.startswith(("displacement_divergence", "mpsa_consistency",
                                                           "boundary_displacement_divergence")):
                continue   # cell-row terms are checked in the cells-mode test
            br = m.shape[0] // g.num_faces
            sel = sps.diags(np.repeat(new, br).astype(float))
            acc[key] = sel @ m if key not in acc else acc[key] + sel @ m
    assert done.all()
    for key, m in acc.items():
        assert rel_err(ref[key], m) < 1e-12, key


def test_cells_mode_rows_and_update_in_place():
    g, flow, mech = problem("cart")
    ref = _full(g, flow, mech)
    cells = np.array([7, 8])
    d1 = _fresh(flow, "flow", specified_cells=cells)
    d2 = _fresh(mech, "mech", specified_cells=cells)
    pb.Mpfa("flow").discretize(g, d1)
    pb.Biot("mech").discretize(g, d2)
    faces = d1[pb.PARAMETERS]["flow"]["active_faces"]
    act_cells = d1[pb.PARAMETERS]["flow"]["active_cells"]
    assert 0 < faces.size < g.num_faces and cells.size < act_cells.size < g.num_cells
    keep_f = np.zeros(g.num_faces, bool)
    keep_f[faces] = True
    part = dict(d1[pb.DISCRETIZATION_MATRICES]["flow"])
    part.update(d2[pb.DISCRETIZATION_MATRICES]["mech"])
    for key, m in flatten(part).items():
        if key.startswith(("displacement_divergence", "mpsa_consistency", "boundary_displacement_divergence")):
            continue
        br = m.shape[0] // g.num_faces
        rows = np.repeat(keep_f, br)
        full = sps.csr_matrix(ref[key])
        assert rel_err(sps.diags(rows.astype(float)) @ full, m) < 1e-12, key      # active rows = full rows
        assert abs(sps.diags((~rows).astype(float)) @ m).sum() == 0, key         # the others are zero
    # update in place: change the permeability in the two cells, re-discretize only around them
    k2 = pb.SecondOrderTensor.from_values(flow[pb.PARAMETERS]["flow"]["second_order_tensor"].values.copy())
    k2.values[:, :, cells] *= 7.0
    flow2 = _fresh(flow, "flow", second_order_tensor=k2)
    pb.Mpfa("flow").discretize(g, flow2)                         # reference result: full pass with k2
    want = flow2[pb.DISCRETIZATION_MATRICES]["flow"]
    upd = _fresh(flow, "flow", second_order_tensor=k2, specified_cells=cells, update_discretization=True)
    upd[pb.DISCRETIZATION_MATRICES]["flow"] = dict(flow[pb.DISCRETIZATION_MATRICES]["flow"])   # old matrices
    pb.Mpfa("flow").discretize(g, upd)
    for key in want:
        assert rel_err(want[key], upd[pb.DISCRETIZATION_MATRICES]["flow"][key]) < 1e-12, key


def test_faces_mode_contains_the_cells_mode_of_the_neighbours():
    g, flow, _ = problem("cart")
    f0 = 40
    d = _fresh(flow, "flow", specified_faces=np.array([f0]))
    cells, faces = fv.active_indices(g, d[pb.PARAMETERS]["flow"])
    assert f0 in faces
    fn = sps.csc_matrix(g.face_nodes)
    nodes = fn.indices[fn.indptr[f0]:fn.indptr[f0 + 1]]
    touching = np.flatnonzero(np.asarray(abs(sps.csr_matrix(g.face_nodes))[nodes].sum(axis=0)).ravel() > 0)
    assert np.array_equal(np.sort(faces), np.sort(touching))
    assert cells.size > 0 and np.all(np.diff(cells) > 0)


def test_biot_update_in_place_incl_cell_row_terms():
    """Stiffness changed in two cells, ``update_discretization = True``: face-row terms of the active
    faces and the cell-row coupling terms of the cells next to them are replaced (biot.py:614-690);
    every stored matrix then equals a full pass with the new stiffness (the sub-grid is grown by one
    ring so that the cell rows that are replaced have complete interaction regions)."""
    # large enough that the two-ring sub-grid of the reference is a proper subset with cut regions:
    # with the reference's row set the neighbours' cell rows come out 1 % wrong on this grid
    g = pb.cart_grid_3d([7, 6, 6], perturb=0.2)
    rng = np.random.default_rng(0)
    nc = g.num_cells
    C = pb.FourthOrderTensor(np.exp(0.3 * rng.standard_normal(nc)), np.exp(0.3 * rng.standard_normal(nc)))
    bf = g.get_all_boundary_faces()
    vbc = pb.BoundaryConditionVectorial(g, bf[g.face_centers[2, bf] < 1e-10], "dir")
    mech = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc,
                                           "scalar_vector_mappings": {"p": 0.8}})
    cells = np.array([3 + 7 * (3 + 6 * 3), 4 + 7 * (3 + 6 * 3)])
    pb.Biot("mech").discretize(g, mech)
    old = mech[pb.DISCRETIZATION_MATRICES]["mech"]
    C2 = pb.FourthOrderTensor.from_values(mech[pb.PARAMETERS]["mech"]["fourth_order_tensor"].values.copy())
    C2.values[:, :, cells] *= 3.0
    full = _fresh(mech, "mech", fourth_order_tensor=C2)
    pb.Biot("mech").discretize(g, full)
    want = full[pb.DISCRETIZATION_MATRICES]["mech"]
    upd = _fresh(mech, "mech", fourth_order_tensor=C2, specified_cells=cells, update_discretization=True)
    upd[pb.DISCRETIZATION_MATRICES]["mech"] = {k: (dict(v) if isinstance(v, dict) else v) for k, v in old.items()}
    pb.Biot("mech").discretize(g, upd)
    got = upd[pb.DISCRETIZATION_MATRICES]["mech"]
    assert upd[pb.PARAMETERS]["mech"]["active_cells"].size < g.num_cells
    for key in ("stress", "bound_stress", "bound_displacement_cell", "bound_displacement_face"):
        assert rel_err(want[key], got[key]) < 1e-12, key
    for key in ("scalar_gradient", "bound_displacement_pressure"):
        for kw in want[key]:
            assert rel_err(want[key][kw], got[key][kw]) < 1e-12, (key, kw)
    # cell-row terms: EVERY row (the neighbours of the modified cells change too, all others must
    # keep their stored values untouched)
    for key in ("displacement_divergence", "boundary_displacement_divergence", "mpsa_consistency"):
        for kw in want[key]:
            assert rel_err(want[key][kw], got[key][kw]) < 1e-12, (key, kw)
            if key != "boundary_displacement_divergence":   # (interior cells: the boundary term does not change)
                assert rel_err(want[key][kw], old[key][kw]) > 1e-6, (key, kw)   # the update was not a no-op


from golden_io import case_names  # noqa: E402
from partial_line_checks import check_partial_update  # noqa: E402


@pytest.mark.parametrize("name", case_names("partial_"))
def test_in_place_update_equals_the_references(name):
    """Host logic of the partial update against golden matrices of the reference's own
    ``update_discretization`` (the GPU leg runs the same check through the CUDA kernels)."""
    check_partial_update(name)
