"""2-D grids embedded in 3-D (fracture planes): the host layer rotates the grid into its own
plane, runs the 2-D kernels and lifts the vector source back to the ambient space
(reference numerics/fv/mpfa.py:733-754, 423-466).  Golden outputs come from the unmodified
reference (tools/make_golden.py, ``embedded_*`` fixtures).

CPU part: ``porepy_b200.fv.Mpfa.discretize`` is driven end to end with the device plan replaced by
the host build of the same node routines (tests/emu) -- this covers all of the host logic that
the feature adds.  The same call on the real plan is tests/test_zz_fracture_planes_gpu.py."""
import numpy as np
import pytest

import porepy_b200 as pb
from porepy_b200 import fv
from cases import max_rel_err
from emu_binding import EmuBackedPlan
from golden_io import case_names, load_case

TOL = 1e-10
CASES = case_names("embedded_")


def _params(c):
    k = pb.SecondOrderTensor.from_values(c.raw["K"])
    return {"second_order_tensor": k, "bc": c.bc, "ambient_dimension": int(c.raw["ambient_dimension"])}


def test_plane_frame_is_a_rotation_into_the_plane():
    c = load_case("embedded_tri2d_tilted")
    R = fv.plane_frame(c.g)
    assert R is not None
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14
    z = (R @ c.g.nodes)[2]
    assert np.ptp(z) < 1e-12
    assert fv.plane_frame(load_case("embedded_cart2d_xy").g) is None
    assert fv.plane_frame(load_case("mpfa_cart2d").g) is None
    bent = load_case("embedded_cart2d_xy").g
    bent.nodes = bent.nodes.copy()
    bent.nodes[2, 0] += 0.05
    with pytest.raises(ValueError):
        fv.plane_frame(bent)


@pytest.mark.parametrize("name", CASES)
def test_fracture_plane_flux_discretization_host_logic(name, monkeypatch):
    c = load_case(name)
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    data = pb.initialize_data({}, "flow", _params(c))
    fv.Mpfa("flow").discretize(c.g, data)
    got = data[pb.DISCRETIZATION_MATRICES]["flow"]
    assert got["vector_source"].shape == (c.g.num_faces, 3 * c.g.num_cells)
    err, key = max_rel_err(c.mats, got)
    assert err < TOL, (key, err)


def test_tilted_plane_without_ambient_dimension(monkeypatch):
    """mpfa.py:459-462: with the default ambient dimension (= 2) the vector-source terms of a tilted plane keep
    the first two ambient components -- the columns (c, 0), (c, 1) of the 3-component golden matrices (the
    reference's own gravity tests run this case, tests/numerics/fv/test_mpfa.py)."""
    c = load_case("embedded_cart2d_tilted")
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    p = _params(c)
    del p["ambient_dimension"]
    data = pb.initialize_data({}, "flow", p)
    fv.Mpfa("flow").discretize(c.g, data)
    nc = c.g.num_cells
    keep = (3 * np.arange(nc)[:, None] + np.arange(2)).ravel()
    for key in ("vector_source", "bound_pressure_vector_source"):
        got, ref = data[pb.DISCRETIZATION_MATRICES]["flow"][key], c.mats[key].tocsc()[:, keep]
        assert got.shape == (c.g.num_faces, 2 * nc)
        assert abs(got - ref).max() <= TOL * abs(ref).max(), key


def test_mechanics_on_a_tilted_plane_is_refused(monkeypatch):
    """The reference's MPSA result on an embedded plane depends on the local frame of map_grid
    (mpsa.py:2005-2040, stiffness not rotated); out of scope, must fail loudly."""
    c = load_case("embedded_cart2d_tilted")
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    nc = c.g.num_cells
    data = pb.initialize_data({}, "mech", {"fourth_order_tensor": pb.FourthOrderTensor(np.ones(nc), np.ones(nc)),
                                           "bc": pb.BoundaryConditionVectorial(c.g)})
    with pytest.raises(NotImplementedError):
        fv.Mpsa("mech").discretize(c.g, data)
