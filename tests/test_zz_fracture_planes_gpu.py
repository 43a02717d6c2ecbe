"""GPU leg of tests/test_embedded_planes.py: flux discretization of fracture planes (2-D grids
embedded in 3-D, ``ambient_dimension = 3``; reference mpfa.py:733-754 / 423-466) through the real
device plan, against golden fixtures written by the reference (3 passed on B200 at the end of round 1)."""
import pytest

import porepy_b200 as pb
from cases import load_case, max_rel_err
from golden_io import case_names

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("name", case_names("embedded_"))
def test_fracture_plane_flux_discretization(name):
    c = load_case(name)
    k = pb.SecondOrderTensor.from_values(c.raw["K"])
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": c.bc,
                                           "ambient_dimension": int(c.raw["ambient_dimension"])})
    pb.Mpfa("flow").discretize(c.g, data)
    got = data[pb.DISCRETIZATION_MATRICES]["flow"]
    assert got["vector_source"].shape == (c.g.num_faces, 3 * c.g.num_cells)
    err, key = max_rel_err(c.mats, got)
    assert err < TOL, (key, err)
