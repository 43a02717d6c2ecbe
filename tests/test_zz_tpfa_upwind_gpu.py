"""GPU leg of tests/test_tpfa_upwind.py: ``pb.Tpfa`` / ``pb.Upwind`` through the real per-face kernels
(``pb_facegrid`` + ``pb_tpfa`` / ``pb_upwind``, one thread per face) against golden outputs of the reference,
incl. a 1-D grid on a tilted line (the TPFA delegation of MPFA / MPSA)."""
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from cases import load_case, max_rel_err
from golden_io import case_names
from partial_line_checks import check_line

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", case_names("next_"))
def test_tpfa_and_upwind_on_the_device(name):
    c = load_case(name)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]),
                                           "bc": c.bc})
    pb.Tpfa("flow").discretize(c.g, data)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("tpfa_")}
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < 1e-13, (key, err)
    td = pb.initialize_data({}, "transport", {"bc": c.bc, "darcy_flux": c.raw["darcy_flux"]})
    pb.Upwind("transport").discretize(c.g, td)
    M = td[pb.DISCRETIZATION_MATRICES]["transport"]
    for ref_key, key in (("upwind", "transport"), ("bound_transport_dir", "rhs_dir"), ("bound_transport_neu", "rhs_neu")):
        assert abs(sps.csr_matrix(c.mats[ref_key]) - M[key]).sum() == 0, key


def test_line_grid_on_the_device():
    check_line()
