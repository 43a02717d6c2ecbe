"""GPU leg of tests/test_tutorial_flux_discretizations.py: the MPFA cell of the reference's tutorial
(``sum(p_mpfa) == 14.192684340967542``) through the CUDA path."""
import numpy as np
import pytest

import porepy_b200 as pb
from test_tutorial_flux_discretizations import solve, tutorial_problem

pytestmark = pytest.mark.gpu


def test_tutorial_mpfa_number_on_the_device():
    g, data = tutorial_problem()
    p = solve(pb.Mpfa("flow"), g, data)
    assert np.isclose(np.sum(p), 14.192684340967542)
