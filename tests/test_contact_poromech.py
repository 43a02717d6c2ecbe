"""Poromechanics of a fractured medium with frictional contact (BASELINE configs[3] + [4] in one model): the reference's
``pp.Poromechanics`` on a matrix cut by a fracture -- Biot poromechanics in the matrix, compressible flow in the fracture
with an aperture that follows the displacement jump, interface Darcy law, fluid pressure on the fracture walls, semismooth
contact laws -- on the device AD chain (porepy_b200/fractured_poromech.py) against the unmodified reference: Jacobian and
residual at the third iterate, the residual history of the semismooth Newton loop and the converged state, for a sliding and
a partly open load case (tests/golden/contact_poromech*.npz, tools/make_contact_golden.py).
CPU: host build of the node / face routines + the scipy stand-in for the device sparse algebra."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200.fractured_poromech import FractureCoupling, FracturedPoromechanics
from porepy_b200.grid import Grid
from golden_io import GOLDEN_DIR

CASES = ["contact_poromech", "contact_poromech_mixed"]


def _csr(d, key):
    return sps.csr_matrix((d[key + "__data"], d[key + "__indices"], d[key + "__indptr"]), shape=tuple(d[key + "__shape"]))


def _grid(d, prefix):
    g = Grid.from_arrays({k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)})
    g.tags["domain_boundary_faces"] = np.asarray(d[prefix + "domain_boundary_faces"], bool)
    if prefix + "tip_faces" in d:
        g.tags["tip_faces"] = np.asarray(d[prefix + "tip_faces"], bool)
    return g


def _flow_bc(d, prefix, nf):
    return SimpleNamespace(is_dir=d[prefix + "flow_is_dir"], is_neu=d[prefix + "flow_is_neu"], is_rob=d[prefix + "flow_is_rob"],
                           is_internal=d[prefix + "flow_is_internal"], robin_weight=np.ones(nf), bc_type="scalar", num_faces=nf)


def load_problem(name):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g, gf = _grid(d, "matrix__"), _grid(d, "fracture__")
    nf = g.num_faces
    vbc = SimpleNamespace(is_dir=d["mech_is_dir"], is_neu=d["mech_is_neu"], is_rob=d["mech_is_rob"],
                          is_internal=d["mech_is_internal"], robin_weight=np.zeros((3, 3, nf)), bc_type="vectorial",
                          num_faces=nf)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(d["matrix__K"]),
                                           "bc": _flow_bc(d, "matrix__", nf)})
    pb.initialize_data(data, "mechanics", {"fourth_order_tensor": pb.FourthOrderTensor.from_values(d["C"]), "bc": vbc,
                                           "scalar_vector_mappings": {"flow": float(d["biot_coefficient"])}})
    fdata = pb.initialize_data({}, "flow", {"bc": _flow_bc(d, "fracture__", gf.num_faces), "ambient_dimension": 3})
    proj = {k: _csr(d, k) for k in ("mortar_to_primary_avg", "primary_to_mortar_int", "mortar_to_secondary_avg",
                                    "secondary_to_mortar_int", "mortar_to_primary_int", "primary_to_mortar_avg",
                                    "mortar_to_secondary_int", "secondary_to_mortar_avg")}
    # the fixture stores the fracture tensor of the initial state: tangential permeability x residual aperture
    frac = FractureCoupling(gf, fdata, proj, d["mortar_sign"], d["mortar_volumes"], _csr(d, "local_coordinates"),
                            d["normal_permeability"], d["fracture__K"] / float(d["residual_aperture"]))
    fluid = {k: float(d[k]) for k in ("compressibility", "density", "viscosity", "reference_pressure")}
    solid = {k: float(d[k]) for k in ("reference_porosity", "n_inv", "residual_aperture")}
    contact = {k: float(d[k]) for k in ("numerical_constant", "characteristic_traction", "friction_coefficient",
                                        "dilation_angle", "reference_gap", "open_state_tolerance")}
    ff = SimpleNamespace(is_dir=d["ff_is_dir"], is_neu=d["ff_is_neu"], is_rob=np.zeros(nf, bool),
                         is_internal=np.asarray(g.tags["fracture_faces"], bool), robin_weight=np.ones(nf), bc_type="scalar",
                         num_faces=nf)
    bc = dict(flow=d["flow_bc_values"], mechanics=d["mech_bc_values"], fluid_flux=d["ff_values"], fluid_flux_type=ff)
    return FracturedPoromechanics(g, data, [frac], fluid, solid, contact, bc), d


def check(prob, d, to_host, make_tensor):
    cm, rm = d["column_map"], d["row_map"]
    assert np.array_equal(np.sort(cm), np.arange(prob.num_dofs)) and np.array_equal(np.sort(rm), np.arange(prob.num_dofs))
    dt = float(d["dt"])
    J, rhs = prob.linearize(d["iterate"][cm], d["previous"][cm], dt)
    Jref = _csr(d, "iterate_jacobian")[rm][:, cm]
    bref = d["iterate_rhs"][rm]
    assert abs(J.to_scipy() - Jref).max() <= 1e-10 * abs(Jref).max()
    assert np.abs(to_host(rhs) - bref).max() <= 1e-10 * max(np.abs(bref).max(), 1e-3 * abs(Jref).max())

    def direct(Jd, r):
        return make_tensor(spla.spsolve(Jd.to_scipy().tocsc(), to_host(r)))
    x, hist = prob.time_step(d["previous"][cm], dt, direct, tol=1e-11)
    ref = d["residual_norms"]
    assert hist[-1]["residual"] <= 1e-10 * hist[0]["residual"] and len(hist) <= len(ref) + 1, hist
    for mine, theirs in zip(hist[:5], ref[:5]):
        if theirs > 1e-9 * ref[0]:
            assert abs(mine["residual"] - theirs) <= 0.05 * theirs, (hist, ref)
    assert np.linalg.norm(to_host(x) - d["solution"][cm]) <= 1e-8 * np.linalg.norm(d["solution"])


@pytest.mark.parametrize("name", CASES)
def test_fractured_poromechanics_with_contact_host_build(name, monkeypatch):
    import torch
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan, emu_interface_upwind_masks
    from porepy_b200 import fv
    import emu_sparse
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    monkeypatch.setattr(fv, "interface_upwind_masks", emu_interface_upwind_masks)
    emu_sparse.install(monkeypatch)
    prob, d = load_problem(name)
    prob.discretize()
    check(prob, d, lambda t: t.numpy(), lambda a: torch.as_tensor(np.asarray(a, float)))
