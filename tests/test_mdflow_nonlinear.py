"""Compressible single-phase flow in a fracture network, the reference's Newton loop restated on the device AD chain
(porepy_b200/mdflow_nl.py): Jacobian and residual at an intermediate iterate and the converged state of one implicit time
step, against the unmodified ``pp.SinglePhaseFlow`` with a compressible fluid (tests/golden/mdflownl_*.npz,
tools/make_mdflow_golden.py ``export_nonlinear``).
CPU: host build of the node / face routines + the scipy stand-in for the device sparse algebra."""
import numpy as np
import pytest

from golden_io import case_names
from mdflow_io import _csr, load_mdflow_nonlinear

CASES = case_names("mdflownl_")


@pytest.fixture()
def host_build(monkeypatch):
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan, emu_interface_upwind_masks
    from porepy_b200 import fv
    import emu_sparse
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    monkeypatch.setattr(fv, "interface_upwind_masks", emu_interface_upwind_masks)
    emu_sparse.install(monkeypatch)


def check_linearization(prob, d, to_host):
    J, rhs = prob.linearize(d["iterate"], d["previous"], float(d["dt"]))
    Jref, bref = _csr(d, "iterate_jacobian"), d["iterate_rhs"]
    assert abs(J.to_scipy() - Jref).max() <= 1e-10 * abs(Jref).max()
    # the iterate is two Newton steps in: the residual is ~1e-5 of the initial one; compare on the Jacobian's scale
    assert np.abs(to_host(rhs) - bref).max() <= 1e-12 * abs(Jref).max()


def check_time_step(prob, d, to_host):
    x, hist = prob.time_step(d["previous"], float(d["dt"]), tol=1e-11)
    assert hist[-1]["residual"] <= 1e-11 * hist[0]["residual"] and len(hist) <= 7, hist
    # quadratic convergence like the reference's own loop (same residual history to two digits)
    ref = d["residual_norms"]
    for mine, theirs in zip(hist[:3], ref[:3]):
        assert abs(mine["residual"] - theirs) <= 0.05 * theirs, (hist, ref)
    assert np.linalg.norm(to_host(x) - d["solution"]) <= 1e-8 * np.linalg.norm(d["solution"])


@pytest.mark.parametrize("name", CASES)
def test_linearization_matches_the_reference_host_build(name, host_build):
    prob, d = load_mdflow_nonlinear(name)
    prob.discretize()
    check_linearization(prob, d, lambda t: t.numpy())


@pytest.mark.parametrize("name", CASES)
def test_time_step_matches_the_reference_host_build(name, host_build):
    prob, d = load_mdflow_nonlinear(name)
    prob.discretize()
    check_time_step(prob, d, lambda t: t.numpy())
