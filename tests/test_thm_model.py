"""``pp.Thermoporomechanics`` (BASELINE config[4] without the fractures): momentum, mass and energy balance, the reference's
model equations on the device AD chain (porepy_b200/thermoporomech.py) against the unmodified reference -- Jacobian and
residual at the third Newton iterate (upwinding taken from that iterate), the residual history and the converged state of one implicit time step (tests/golden/thm_model.npz,
tools/make_thm_golden.py).
CPU: host build of the node / face routines + the scipy stand-in for the device sparse algebra."""
import os
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from porepy_b200.grid import Grid
from porepy_b200.thermoporomech import Thermoporomechanics
from golden_io import GOLDEN_DIR


def _scalar_bc(d, prefix, nf):
    return SimpleNamespace(is_dir=d[prefix + "_is_dir"], is_neu=d[prefix + "_is_neu"], is_rob=np.zeros(nf, bool),
                           is_internal=np.zeros(nf, bool), robin_weight=np.ones(nf), bc_type="scalar", num_faces=nf)


def load_problem():
    d = dict(np.load(os.path.join(GOLDEN_DIR, "thm_model.npz"), allow_pickle=False))
    g = Grid.from_arrays(d)
    nf = g.num_faces
    vbc = SimpleNamespace(is_dir=d["mech_is_dir"], is_neu=d["mech_is_neu"], is_rob=d["mech_is_rob"],
                          is_internal=d["mech_is_internal"], robin_weight=np.zeros((3, 3, nf)), bc_type="vectorial",
                          num_faces=nf)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(d["K"]),
                                           "bc": _scalar_bc(d, "flow", nf)})
    pb.initialize_data(data, "fourier", {"bc": _scalar_bc(d, "fourier", nf)})
    pb.initialize_data(data, "mechanics", {
        "fourth_order_tensor": pb.FourthOrderTensor.from_values(d["C"]), "bc": vbc,
        "scalar_vector_mappings": {"flow": pb.SecondOrderTensor.from_values(d["alpha_flow"]),
                                   "thermal": pb.SecondOrderTensor.from_values(d["alpha_thermal"])}})
    fluid = dict(compressibility=d["compressibility"], density=d["density"], viscosity=d["viscosity"],
                 thermal_expansion=d["fluid_thermal_expansion"], heat_capacity=d["fluid_heat_capacity"],
                 conductivity=d["fluid_conductivity"], reference_pressure=d["reference_pressure"],
                 reference_temperature=d["reference_temperature"])
    solid = dict(reference_porosity=d["reference_porosity"], n_inv=d["n_inv"], biot_coefficient=d["biot_coefficient"],
                 thermal_expansion=d["solid_thermal_expansion"], heat_capacity=d["solid_heat_capacity"],
                 conductivity=d["solid_conductivity"], density=d["solid_density"])
    bc = dict(flow=d["flow_bc_values"], fourier=d["fourier_bc_values"], mechanics=d["mech_bc_values"],
              fluid_flux=d["ff_values"], enthalpy_flux=d["ef_values"], fluid_flux_type=_scalar_bc(d, "ff", nf),
              enthalpy_flux_type=_scalar_bc(d, "ef", nf))
    return Thermoporomechanics(g, data, fluid, solid, bc), d


def _csr(d, key):
    return sps.csr_matrix((d[key + "__data"], d[key + "__indices"], d[key + "__indptr"]), shape=tuple(d[key + "__shape"]))


def check(prob, d, to_host, linear_solver=None):
    J, rhs = prob.linearize(d["iterate"], d["previous"], float(d["dt"]))
    Jref, bref = _csr(d, "iterate_jacobian"), d["iterate_rhs"]
    assert abs(J.to_scipy() - Jref).max() <= 1e-10 * abs(Jref).max()
    assert np.abs(to_host(rhs) - bref).max() <= 1e-10 * np.abs(bref).max()
    x, hist = prob.time_step(d["previous"], float(d["dt"]), tol=1e-11, linear_solver=linear_solver)
    ref = d["residual_norms"]
    assert hist[-1]["residual"] <= 1e-11 * hist[0]["residual"] and len(hist) <= len(ref) + 1, hist
    for mine, theirs in zip(hist[:4], ref[:4]):            # the reference's own convergence, step by step
        assert abs(mine["residual"] - theirs) <= 0.05 * theirs, (hist, ref)
    assert np.linalg.norm(to_host(x) - d["solution"]) <= 1e-8 * np.linalg.norm(d["solution"])
    return hist


def test_thermoporomechanics_model_host_build(monkeypatch):
    import torch
    from emu_binding import EmuBackedFaceGrid, EmuBackedPlan
    from porepy_b200 import fv
    import emu_sparse
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    emu_sparse.install(monkeypatch)
    prob, d = load_problem()
    prob.discretize()

    def direct(J, rhs):
        return torch.as_tensor(spla.spsolve(J.to_scipy().tocsc(), rhs.numpy()))
    check(prob, d, lambda t: t.numpy(), linear_solver=direct)
