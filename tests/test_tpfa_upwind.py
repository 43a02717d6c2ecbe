"""Two-point flux approximation and first-order upwinding (reference
numerics/fv/tpfa.py, upwind.py): the per-face routines of csrc/face_kernels.cuh (host build) and the
operator classes ``pb.Tpfa`` / ``pb.Upwind`` driven through the emulation-backed plan, against golden
outputs of the reference (``next_*`` fixtures) and against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200 import fv
from cases import load_case, max_rel_err
from emu_binding import EmuBackedFaceGrid, EmuBackedPlan, EmuPlan
from golden_io import case_names

CASES = case_names("next_")


@pytest.mark.parametrize("name", CASES)
def test_face_routines_match_the_reference(name):
    c = load_case(name)
    p = EmuPlan(c.g)
    out = p.tpfa(c.raw["K"], c.bc)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("tpfa_")}
    err, key = max_rel_err(ref, out)
    assert err < 1e-14, (key, err)
    up = p.upwind(c.raw["darcy_flux"], c.bc)
    for key in ("upwind", "bound_transport_dir", "bound_transport_neu"):
        assert abs(sps.csr_matrix(c.mats[key]) - up[key]).sum() == 0, key


@pytest.mark.parametrize("name", CASES)
def test_operator_classes(name, monkeypatch):
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    c = load_case(name)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]),
                                           "bc": c.bc})
    d = pb.Tpfa("flow")
    d.discretize(c.g, data)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("tpfa_")}
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < 1e-14, (key, err)
    # TPFA is what MPFA reduces to on an orthogonal grid with isotropic K: same assembled operator class
    data[pb.PARAMETERS]["flow"]["bc_values"] = np.zeros(c.g.num_faces)
    A, b = d.assemble_matrix_rhs(c.g, data)
    assert A.shape == (c.g.num_cells, c.g.num_cells) and b.shape == (c.g.num_cells,)
    u = pb.Upwind("transport")
    td = pb.initialize_data({}, "transport", {"bc": c.bc, "darcy_flux": c.raw["darcy_flux"],
                                              "bc_values": np.ones(c.g.num_faces)})
    u.discretize(c.g, td)
    M = td[pb.DISCRETIZATION_MATRICES]["transport"]
    assert abs(sps.csr_matrix(c.mats["upwind"]) - M["transport"]).sum() == 0
    assert abs(sps.csr_matrix(c.mats["bound_transport_dir"]) - M["rhs_dir"]).sum() == 0
    assert abs(sps.csr_matrix(c.mats["bound_transport_neu"]) - M["rhs_neu"]).sum() == 0
    A, b = u.assemble_matrix_rhs(c.g, td)
    assert A.shape == (c.g.num_cells, c.g.num_cells)
    # conservation: without boundaries every column of div @ diag(q) @ upwind sums to the net outflow
    td["parameters"]["transport"]["num_components"] = 2
    u.discretize(c.g, td)
    assert td[pb.DISCRETIZATION_MATRICES]["transport"]["transport"].shape == (2 * c.g.num_faces, 2 * c.g.num_cells)


def test_line_grid_delegation(monkeypatch):
    """1-D grid on a tilted line in 3-D: MPFA / MPSA delegate to TPFA (mpfa.py:690-712, mpsa.py:666-697)."""
    from partial_line_checks import check_line
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    check_line()
