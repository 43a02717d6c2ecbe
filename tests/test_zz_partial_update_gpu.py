"""GPU leg of tests/test_partial_update.py: partial (re)discretization around two cells through the
real device plan (sub-grid extraction and row embedding are host logic, the sub-grid runs the same
kernels): in-place update of MPFA and Biot equals a full pass with the new parameters."""
import numpy as np
import pytest

import porepy_b200 as pb
from golden_io import case_names, rel_err
from partial_line_checks import check_partial_update

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", case_names("partial_"))
def test_in_place_update_equals_the_references(name):
    """Against the reference itself: golden matrices of ``Mpfa / Mpsa.update_discretization`` after two cells
    changed (tools/make_golden.py: case_partial_update)."""
    check_partial_update(name)


def test_update_in_place_equals_a_full_pass():
    g = pb.cart_grid_3d([7, 6, 6], perturb=0.2)
    rng = np.random.default_rng(0)
    nc = g.num_cells
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    C = pb.FourthOrderTensor(np.exp(0.3 * rng.standard_normal(nc)), np.exp(0.3 * rng.standard_normal(nc)))
    bf = g.get_all_boundary_faces()
    bc = pb.BoundaryCondition(g, bf[g.face_centers[0, bf] < 1e-10], "dir")
    vbc = pb.BoundaryConditionVectorial(g, bf[g.face_centers[2, bf] < 1e-10], "dir")
    cells = np.array([3 + 7 * (3 + 6 * 3), 4 + 7 * (3 + 6 * 3)])

    def flow(kk, **extra):
        return pb.initialize_data({}, "flow", {"second_order_tensor": kk, "bc": bc, **extra})

    def mech(cc, **extra):
        return pb.initialize_data({}, "mech", {"fourth_order_tensor": cc, "bc": vbc,
                                               "scalar_vector_mappings": {"p": 0.8}, **extra})
    old_f, old_m = flow(k), mech(C)
    pb.Mpfa("flow").discretize(g, old_f)
    pb.Biot("mech").discretize(g, old_m)
    k2 = pb.SecondOrderTensor.from_values(k.values.copy())
    k2.values[:, :, cells] *= 7.0
    C2 = pb.FourthOrderTensor.from_values(C.values.copy())
    C2.values[:, :, cells] *= 3.0
    want_f, want_m = flow(k2), mech(C2)
    pb.Mpfa("flow").discretize(g, want_f)
    pb.Biot("mech").discretize(g, want_m)
    upd_f = flow(k2, specified_cells=cells, update_discretization=True)
    upd_f[pb.DISCRETIZATION_MATRICES]["flow"] = dict(old_f[pb.DISCRETIZATION_MATRICES]["flow"])
    pb.Mpfa("flow").discretize(g, upd_f)
    assert upd_f[pb.PARAMETERS]["flow"]["active_cells"].size < nc
    for key, m in want_f[pb.DISCRETIZATION_MATRICES]["flow"].items():
        assert rel_err(m, upd_f[pb.DISCRETIZATION_MATRICES]["flow"][key]) < 1e-12, key
    upd_m = mech(C2, specified_cells=cells, update_discretization=True)
    upd_m[pb.DISCRETIZATION_MATRICES]["mech"] = {
        key: (dict(v) if isinstance(v, dict) else v) for key, v in old_m[pb.DISCRETIZATION_MATRICES]["mech"].items()}
    pb.Biot("mech").discretize(g, upd_m)
    for key, m in want_m[pb.DISCRETIZATION_MATRICES]["mech"].items():
        got = upd_m[pb.DISCRETIZATION_MATRICES]["mech"][key]
        if isinstance(m, dict):
            for kw in m:
                assert rel_err(m[kw], got[kw]) < 1e-12, (key, kw)
        else:
            assert rel_err(m, got) < 1e-12, key
