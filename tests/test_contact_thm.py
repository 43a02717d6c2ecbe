"""Thermo-poromechanics of a fractured medium with frictional contact -- BASELINE config[4] as the reference states it
(``pp.Thermoporomechanics`` on a matrix cut by a fracture; 388 unknowns in 10 variable groups, 11 equation groups) -- on
the device AD chain (porepy_b200/fractured_thm.py) against the unmodified reference: the Jacobian at the zero state (every
tie rule of the semismooth laws and of the upwinding active) and at the fourth Newton iterate (aperture off its residual
value: re-discretized fracture fluxes), the residual history of the semismooth Newton loop and the converged state, for a
sliding and a partly open load case (tests/golden/contact_thm*.npz, tools/make_contact_golden.py).
CPU: host build of the node / face routines + the scipy stand-in for the device sparse algebra."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200.fractured_poromech import FractureCoupling
from porepy_b200.fractured_thm import FracturedThermoporomechanics
from test_contact_poromech import _csr, _grid
from golden_io import GOLDEN_DIR

CASES = ["contact_thm", "contact_thm_mixed"]


def _bc(d, prefix, nf, internal=None):
    return SimpleNamespace(is_dir=d[prefix + "is_dir"], is_neu=d[prefix + "is_neu"],
                           is_rob=d.get(prefix + "is_rob", np.zeros(nf, bool)),
                           is_internal=d.get(prefix + "is_internal", internal if internal is not None else np.zeros(nf, bool)),
                           robin_weight=np.ones(nf), bc_type="scalar", num_faces=nf)


def load_problem(name):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g, gf = _grid(d, "matrix__"), _grid(d, "fracture__")
    nf, nff = g.num_faces, gf.num_faces
    vbc = SimpleNamespace(is_dir=d["mech_is_dir"], is_neu=d["mech_is_neu"], is_rob=d["mech_is_rob"],
                          is_internal=d["mech_is_internal"], robin_weight=np.zeros((3, 3, nf)), bc_type="vectorial",
                          num_faces=nf)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(d["matrix__flow_K"]),
                                           "bc": _bc(d, "matrix__flow_", nf)})
    pb.initialize_data(data, "fourier", {"bc": _bc(d, "matrix__fourier_", nf)})
    pb.initialize_data(data, "mechanics", {
        "fourth_order_tensor": pb.FourthOrderTensor.from_values(d["C"]), "bc": vbc,
        "scalar_vector_mappings": {"flow": pb.SecondOrderTensor.from_values(d["alpha_flow"]),
                                   "thermal": pb.SecondOrderTensor.from_values(d["alpha_thermal"])}})
    fdata = pb.initialize_data({}, "flow", {"bc": _bc(d, "fracture__flow_", nff), "ambient_dimension": 3})
    pb.initialize_data(fdata, "fourier", {"bc": _bc(d, "fracture__fourier_", nff), "ambient_dimension": 3})
    proj = {k: _csr(d, k) for k in ("mortar_to_primary_avg", "primary_to_mortar_int", "mortar_to_secondary_avg",
                                    "secondary_to_mortar_int", "mortar_to_primary_int", "primary_to_mortar_avg",
                                    "mortar_to_secondary_int", "secondary_to_mortar_avg")}
    frac = FractureCoupling(gf, fdata, proj, d["mortar_sign"], d["mortar_volumes"], _csr(d, "local_coordinates"),
                            d["normal_permeability"], d["fracture__flow_K"] / float(d["residual_aperture"]))
    fluid = dict(compressibility=d["compressibility"], density=d["density"], viscosity=d["viscosity"],
                 reference_pressure=d["reference_pressure"], thermal_expansion=d["fluid_thermal_expansion"],
                 heat_capacity=d["fluid_heat_capacity"], conductivity=d["fluid_conductivity"],
                 reference_temperature=d["reference_temperature"])
    solid = dict(reference_porosity=d["reference_porosity"], n_inv=d["n_inv"], residual_aperture=d["residual_aperture"],
                 biot_coefficient=d["biot_coefficient"], thermal_expansion=d["solid_thermal_expansion"],
                 heat_capacity=d["solid_heat_capacity"], conductivity=d["solid_conductivity"], density=d["solid_density"])
    contact = {k: float(d[k]) for k in ("numerical_constant", "characteristic_traction", "friction_coefficient",
                                        "dilation_angle", "reference_gap", "open_state_tolerance")}
    internal = np.asarray(g.tags["fracture_faces"], bool)
    bc = dict(flow=d["flow_bc_values"], fourier=d["fourier_bc_values"], mechanics=d["mech_bc_values"],
              fluid_flux=d["ff_values"], enthalpy_flux=d["ef_values"], fluid_flux_type=_bc(d, "ff_", nf, internal),
              enthalpy_flux_type=_bc(d, "ef_", nf, internal))
    prob = FracturedThermoporomechanics(g, data, [frac], fluid, solid, contact, bc, [d["normal_thermal_conductivity"]])
    return prob, d


def check(prob, d, to_host, make_tensor):
    cm, rm = d["column_map"], d["row_map"]
    assert np.array_equal(np.sort(cm), np.arange(prob.num_dofs)) and np.array_equal(np.sort(rm), np.arange(prob.num_dofs))
    dt = float(d["dt"])
    for state, jac, rhs_key in ((d["previous"], "initial_jacobian", "initial_rhs"), (d["iterate"], "iterate_jacobian", "iterate_rhs")):
        J, rhs = prob.linearize(state[cm], d["previous"][cm], dt)
        Jref, bref = _csr(d, jac)[rm][:, cm], d[rhs_key][rm]
        assert abs(J.to_scipy() - Jref).max() <= 1e-10 * abs(Jref).max(), jac
        assert np.abs(to_host(rhs) - bref).max() <= 1e-10 * max(np.abs(bref).max(), 1e-3 * abs(Jref).max()), rhs_key

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
def test_fractured_thermoporomechanics_with_contact_host_build(name, monkeypatch):
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
