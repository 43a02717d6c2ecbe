"""GPU leg of tests/test_tpfa_upwind.py: ``pb.Tpfa`` / ``pb.Upwind`` through the real device plan
(``pb_tpfa`` / ``pb_upwind``, one thread per face) against golden outputs of the reference.

These kernels were added after the round's GPU budget was spent: the per-face routines are validated
through the host build (bitwise equal to the reference) and the library cross-compiles, but the CUDA
launch path has not been executed on a B200 yet -- hence non-strict xfail (an XPASS is the expected
outcome; a failure here does not touch the validated MPFA / MPSA paths, whose SASS is unchanged)."""
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from cases import load_case, max_rel_err
from golden_io import case_names

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first run of pb_tpfa / pb_upwind on a GPU")]


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
