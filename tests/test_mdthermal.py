"""Mass and energy balance in a fracture network (``pp.MassAndEnergyBalance``; the thermal half of BASELINE config[4] on a
mixed-dimensional grid) on the device AD chain (porepy_b200/mdthermal.py) against the unmodified reference: Jacobian and
residual at the third Newton iterate, residual history and converged state of one implicit time step
(tests/golden/mdthermal_*.npz; the fixtures carry the maps between the reference's interleaved numbering and
[p | T | lambda | eta | eps]).  CPU: host build + the scipy stand-in for the device sparse algebra."""
import numpy as np
import pytest

from golden_io import case_names
from mdflow_io import _csr, load_mdthermal

CASES = case_names("mdthermal_")


@pytest.fixture()
def host_build(monkeypatch):
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan, emu_interface_upwind_masks
    from porepy_b200 import fv
    import emu_sparse
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    monkeypatch.setattr(fv, "interface_upwind_masks", emu_interface_upwind_masks)
    emu_sparse.install(monkeypatch)


def check(prob, d, to_host):
    cm, rm = d["column_map"], d["row_map"]
    assert np.array_equal(np.sort(cm), np.arange(prob.num_dofs)) and np.array_equal(np.sort(rm), np.arange(prob.num_dofs))
    J, rhs = prob.linearize(d["iterate"][cm], d["previous"][cm], float(d["dt"]))
    Jref = _csr(d, "iterate_jacobian")[rm][:, cm]
    bref = d["iterate_rhs"][rm]
    assert abs(J.to_scipy() - Jref).max() <= 1e-10 * abs(Jref).max()
    assert np.abs(to_host(rhs) - bref).max() <= 1e-10 * np.abs(bref).max()
    x, hist = prob.time_step(d["previous"][cm], float(d["dt"]), tol=1e-11)
    ref = d["residual_norms"]
    assert hist[-1]["residual"] <= 1e-11 * hist[0]["residual"] and len(hist) <= len(ref) + 1, hist
    for mine, theirs in zip(hist[:4], ref[:4]):
        assert abs(mine["residual"] - theirs) <= 0.05 * theirs, (hist, ref)
    assert all(h.get("linear_converged", True) for h in hist), hist
    assert np.linalg.norm(to_host(x) - d["solution"][cm]) <= 1e-8 * np.linalg.norm(d["solution"])


@pytest.mark.parametrize("name", CASES)
def test_mass_and_energy_balance_on_a_network_host_build(name, host_build):
    prob, d = load_mdthermal(name)
    prob.discretize()
    check(prob, d, lambda t: t.numpy())
