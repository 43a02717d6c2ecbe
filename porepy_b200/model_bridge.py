"""From a prepared PorePy model to the device problems of this package: the glue a PorePy user needs to hand a live
``pp.SinglePhaseFlow`` / ``pp.MassAndEnergyBalance`` (on a fracture network) or ``pp.Poromechanics`` /
``pp.Thermoporomechanics`` (3-D subdomain) over to ``porepy_b200`` -- grids, parameter dictionaries and mortar projections
of ``model.mdg`` as they are; boundary data, coefficients and constants evaluated from the model's own methods
(``bc_values_*``, ``bc_type_*``, ``normal_permeability``, ``aperture``, ``specific_volume``, ``porosity``, the fluid and
solid constants).  Reached through ``porepy_plugin.plugin(pp)``: ``b200.compressible_flow_from_model(model)`` etc.

The model's data dictionaries are not modified: every problem gets its own dictionaries (the parameter entries are shared,
the discretization matrices and the upwind parameters are the problem's own).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

from .params import DISCRETIZATION_MATRICES, PARAMETERS


def _evaluated(model, op, n):
    v = model.equation_system.evaluate(op)
    v = getattr(v, "val", v)
    return np.full(n, float(v)) if np.ndim(v) == 0 else np.asarray(v, float)


def _own_data(data: dict, keywords) -> dict:
    return {PARAMETERS: {kw: dict(data[PARAMETERS][kw]) for kw in keywords if kw in data.get(PARAMETERS, {})},
            DISCRETIZATION_MATRICES: {}}


def _face_values(model, sd, bc, dirichlet, neumann):
    """Boundary operator of a flux law on ``sd``: ``dirichlet(bg)`` on the Dirichlet faces of ``bc``, ``neumann(bg)``
    elsewhere (``_combine_boundary_operators`` of the models), as a face array."""
    bg = model.mdg.subdomain_to_boundary_grid(sd)
    if bg is None or bg.num_cells == 0:
        return np.zeros(sd.num_faces)
    proj = bg.projection()
    return np.where(bc.is_dir, proj.T @ dirichlet(bg), proj.T @ neumann(bg))


def _fluid(model, thermal: bool) -> dict:
    fl = model.fluid.reference_component
    out = dict(compressibility=fl.compressibility, density=fl.density, viscosity=fl.viscosity,
               reference_pressure=model.reference_variable_values.pressure)
    if thermal:
        out.update(thermal_expansion=fl.thermal_expansion, heat_capacity=fl.specific_heat_capacity,
                   conductivity=fl.thermal_conductivity, reference_temperature=model.reference_variable_values.temperature)
    return out


def _boundary_weights(model, sd, fluid: dict, thermal: bool):
    """(rho / mu, c_f (T - T0) rho / mu) of the boundary pressure / temperature, as functions of a boundary grid."""
    def rho(bg):
        e = fluid["compressibility"] * (model.bc_values_pressure(bg) - fluid["reference_pressure"])
        if thermal:
            e = e - fluid["thermal_expansion"] * (model.bc_values_temperature(bg) - fluid["reference_temperature"])
        return fluid["density"] * np.exp(e)

    def w(bg):
        return rho(bg) / fluid["viscosity"]

    def we(bg):
        return fluid["heat_capacity"] * (model.bc_values_temperature(bg) - fluid["reference_temperature"]) * w(bg)
    return w, we


def _interfaces(model, thermal: bool):
    from .mdflow import MdInterface
    mdg = model.mdg
    index = {id(sd): i for i, sd in enumerate(mdg.subdomains())}
    out, kappa_t = [], []
    for it in mdg.interfaces():
        if getattr(it, "codim", 1) != 1:
            continue
        h, l = mdg.interface_to_subdomain_pair(it)
        out.append(MdInterface(index[id(h)], index[id(l)], it.mortar_to_primary_int(), it.primary_to_mortar_avg(),
                               it.mortar_to_secondary_int(), it.secondary_to_mortar_avg(),
                               _evaluated(model, model.normal_permeability([it]), it.num_cells),
                               np.asarray(it.cell_volumes, float) * _evaluated(model, model.specific_volume([it]), it.num_cells),
                               _evaluated(model, model.aperture([l]), l.num_cells)))
        if thermal:
            kappa_t.append(_evaluated(model, model.normal_thermal_conductivity([it]), it.num_cells))
    return out, kappa_t


def compressible_flow_from_model(model):
    """``pp.SinglePhaseFlow`` (compressible fluid) on ``model.mdg`` -> ``CompressibleMixedDimensionalFlow``; unknowns and
    equations in the model's own order."""
    from .mdflow import MdSubdomain
    from .mdflow_nl import CompressibleMixedDimensionalFlow
    mdg, kw = model.mdg, model.darcy_keyword
    fluid = _fluid(model, False)
    subs, storage, bcs, weights = [], [], [], []
    for sd in mdg.subdomains():
        data = _own_data(mdg.subdomain_data(sd), [kw])
        n = sd.num_cells
        storage.append(np.asarray(sd.cell_volumes, float) * _evaluated(model, model.specific_volume([sd]), n)
                       * _evaluated(model, model.porosity([sd]), n))
        if sd.num_faces == 0:
            subs.append(MdSubdomain(sd, data))
            bcs.append(None)
            weights.append(None)
            continue
        w, _ = _boundary_weights(model, sd, fluid, False)
        bc_ff = model.bc_type_fluid_flux(sd)
        subs.append(MdSubdomain(sd, data, _face_values(model, sd, data[PARAMETERS][kw]["bc"], model.bc_values_pressure,
                                                       model.bc_values_darcy_flux)))
        bcs.append(bc_ff)
        weights.append(_face_values(model, sd, bc_ff, w, model.bc_values_fluid_flux))
    intfs, _ = _interfaces(model, False)
    prob = CompressibleMixedDimensionalFlow(subs, intfs, fluid, storage, bcs, weights, keyword=kw)
    prob.mobility_keyword = "b200_mobility"
    return prob


def mass_energy_from_model(model):
    """``pp.MassAndEnergyBalance`` on ``model.mdg`` -> (``MixedDimensionalMassEnergy``, column_map, row_map): unknown k
    of the problem is dof ``column_map[k]`` of the model's ``EquationSystem``, equation k its row ``row_map[k]``."""
    from .mdflow import MdSubdomain
    from .mdthermal import MixedDimensionalMassEnergy
    mdg, es = model.mdg, model.equation_system
    fk, tk = model.darcy_keyword, model.fourier_keyword
    fluid = _fluid(model, True)
    solid = dict(density=model.solid.density, heat_capacity=model.solid.specific_heat_capacity)
    sds = list(mdg.subdomains())
    subs, volume, porosity, bcv, bct = [], [], [], [], []
    for sd in sds:
        data = _own_data(mdg.subdomain_data(sd), [fk, tk])
        n = sd.num_cells
        volume.append(np.asarray(sd.cell_volumes, float) * _evaluated(model, model.specific_volume([sd]), n))
        porosity.append(_evaluated(model, model.porosity([sd]), n))
        subs.append(MdSubdomain(sd, data))
        if sd.num_faces == 0:
            bcv.append(None)
            bct.append(None)
            continue
        w, we = _boundary_weights(model, sd, fluid, True)
        ff, ef = model.bc_type_fluid_flux(sd), model.bc_type_enthalpy_flux(sd)
        prm = data[PARAMETERS]
        bcv.append(dict(flow=_face_values(model, sd, prm[fk]["bc"], model.bc_values_pressure, model.bc_values_darcy_flux),
                        fourier=_face_values(model, sd, prm[tk]["bc"], model.bc_values_temperature, model.bc_values_fourier_flux),
                        fluid_flux=_face_values(model, sd, ff, w, model.bc_values_fluid_flux),
                        enthalpy_flux=_face_values(model, sd, ef, we, model.bc_values_enthalpy_flux)))
        bct.append(dict(fluid_flux=ff, enthalpy_flux=ef))
    intfs, kappa_t = _interfaces(model, True)
    prob = MixedDimensionalMassEnergy(subs, intfs, fluid, solid, volume, porosity, bcv, bct, kappa_t, flow_keyword=fk,
                                      fourier_keyword=tk)
    prob.mobility_keyword, prob.enthalpy_upwind_keyword = "b200_mobility", "b200_enthalpy_upwind"
    its = [it for it in mdg.interfaces() if getattr(it, "codim", 1) == 1]

    def dofs(name, g):
        return es.dofs_of([v for v in es.variables if v.name == name and v.domain is g])
    cols = [dofs(name, g) for name, grids in ((model.pressure_variable, sds), (model.temperature_variable, sds),
                                               (model.interface_darcy_flux_variable, its),
                                               (model.interface_fourier_flux_variable, its),
                                               (model.interface_enthalpy_flux_variable, its)) for g in grids]
    rows, r0 = {}, 0
    for eq in es.equations:
        grids = sds if eq in ("mass_balance_equation", "energy_balance_equation") else its if eq.startswith("interface") else []
        for g in grids:
            rows[(eq, id(g))] = np.arange(r0, r0 + g.num_cells)
            r0 += g.num_cells
    order = [("mass_balance_equation", sds), ("energy_balance_equation", sds), ("interface_darcy_flux_equation", its),
             ("interface_fourier_flux_equation", its), ("interface_enthalpy_flux_equation", its)]
    return prob, np.concatenate(cols), np.concatenate([rows[(eq, id(g))] for eq, grids in order for g in grids])


def _mechanics_boundary(model, sd, data, mk):
    bg = model.mdg.subdomain_to_boundary_grid(sd)
    proj3 = sps.kron(bg.projection(), sps.identity(3)).tocsr()
    bc = data[PARAMETERS][mk]["bc"]
    return np.where(np.asarray(bc.is_dir).ravel("F"), proj3.T @ model.bc_values_displacement(bg),
                    proj3.T @ model.bc_values_stress(bg))


def _n_inv(model):
    so = model.solid
    bulk = so.lame_lambda + 2.0 * so.shear_modulus / 3.0          # ``bulk_modulus`` of the solid constants
    return (so.biot_coefficient - so.porosity) * (1.0 - so.biot_coefficient) / bulk


def poromechanics_from_model(model):
    """``pp.Poromechanics`` on a 3-D subdomain without fractures -> ``Poromechanics`` (unknowns [p | u], the model's
    own order)."""
    from .poromech import Poromechanics
    sds = list(model.mdg.subdomains())
    if len(sds) != 1 or sds[0].dim != 3:
        raise NotImplementedError("one 3-D subdomain without fractures is expected")
    sd = sds[0]
    fk, mk = model.darcy_keyword, model.stress_keyword
    data = _own_data(model.mdg.subdomain_data(sd), [fk, mk])
    fluid = _fluid(model, False)
    w, _ = _boundary_weights(model, sd, fluid, False)
    bc_ff = model.bc_type_fluid_flux(sd)
    prob = Poromechanics(sd, data, fluid, dict(reference_porosity=model.solid.porosity, n_inv=_n_inv(model)),
                         _face_values(model, sd, data[PARAMETERS][fk]["bc"], model.bc_values_pressure, model.bc_values_darcy_flux),
                         _mechanics_boundary(model, sd, data, mk), bc_ff,
                         _face_values(model, sd, bc_ff, w, model.bc_values_fluid_flux), flow_keyword=fk, mechanics_keyword=mk)
    prob.mobility_keyword = "b200_mobility"
    return prob


def thermoporomechanics_from_model(model):
    """``pp.Thermoporomechanics`` on a 3-D subdomain without fractures -> ``Thermoporomechanics`` (unknowns [u | p | T],
    the model's own order)."""
    from .thermoporomech import Thermoporomechanics
    sds = list(model.mdg.subdomains())
    if len(sds) != 1 or sds[0].dim != 3:
        raise NotImplementedError("one 3-D subdomain without fractures is expected")
    sd = sds[0]
    fk, tk, mk, ck = model.darcy_keyword, model.fourier_keyword, model.stress_keyword, model.enthalpy_keyword
    data = _own_data(model.mdg.subdomain_data(sd), [fk, tk, mk])
    fluid = _fluid(model, True)
    so = model.solid
    solid = dict(reference_porosity=so.porosity, n_inv=_n_inv(model), biot_coefficient=so.biot_coefficient,
                 thermal_expansion=so.thermal_expansion, heat_capacity=so.specific_heat_capacity,
                 conductivity=so.thermal_conductivity, density=so.density)
    w, we = _boundary_weights(model, sd, fluid, True)
    ff, ef = model.bc_type_fluid_flux(sd), model.bc_type_enthalpy_flux(sd)
    prm = data[PARAMETERS]
    bc = dict(flow=_face_values(model, sd, prm[fk]["bc"], model.bc_values_pressure, model.bc_values_darcy_flux),
              fourier=_face_values(model, sd, prm[tk]["bc"], model.bc_values_temperature, model.bc_values_fourier_flux),
              mechanics=_mechanics_boundary(model, sd, data, mk),
              fluid_flux=_face_values(model, sd, ff, w, model.bc_values_fluid_flux),
              enthalpy_flux=_face_values(model, sd, ef, we, model.bc_values_enthalpy_flux),
              fluid_flux_type=ff, enthalpy_flux_type=ef)
    prob = Thermoporomechanics(sd, data, fluid, solid, bc, flow_keyword=fk, fourier_keyword=tk, mechanics_keyword=mk,
                               thermal_keyword=ck)
    prob.mobility_keyword, prob.enthalpy_upwind_keyword = "b200_mobility", "b200_enthalpy_upwind"
    return prob


def fractured_momentum_from_model(model):
    """``pp.MomentumBalance`` with fractures in frictional contact -> (``FracturedMomentumBalance``, column_map): one 3-D
    matrix subdomain, any number of 2-D fractures (each with its two-sided interface); unknown k of the problem is dof
    ``column_map[k]`` of the model ([u | contact tractions | interface displacements])."""
    from .contact import FractureContact, FracturedMomentumBalance
    mdg, es = model.mdg, model.equation_system
    mats = list(mdg.subdomains(dim=3))
    fracs = list(mdg.subdomains(dim=2))
    if len(mats) != 1 or any(sd.dim < 2 for sd in mdg.subdomains()):
        raise NotImplementedError("one 3-D matrix subdomain and 2-D fractures without intersections are expected")
    mat = mats[0]
    mk = model.stress_keyword
    data = _own_data(mdg.subdomain_data(mat), [mk])

    def scalar(op):
        return float(np.atleast_1d(_evaluated(model, op, 1))[0])
    contacts, intfs = [], []
    for frac in fracs:
        intf = [it for it in mdg.interfaces() if mdg.interface_to_subdomain_pair(it)[1] is frac][0]
        rot = mdg.subdomain_data(frac)["tangential_normal_projection"].project_tangential_normal(frac.num_cells)
        contacts.append(FractureContact(intf.mortar_to_primary_avg(), intf.primary_to_mortar_int(),
                                        intf.mortar_to_secondary_avg(), intf.secondary_to_mortar_int(),
                                        sps.csr_matrix(intf.sign_of_mortar_sides(1)).diagonal(), intf.cell_volumes, rot))
        intfs.append(intf)
    constants = dict(numerical_constant=scalar(model.contact_mechanics_numerical_constant(fracs)),
                     characteristic_traction=scalar(model.characteristic_contact_traction(fracs)),
                     friction_coefficient=scalar(model.friction_coefficient(fracs)),
                     dilation_angle=model.solid.dilation_angle, reference_gap=model.solid.fracture_gap,
                     open_state_tolerance=model.numerical.open_state_tolerance)
    prob = FracturedMomentumBalance(mat, data, _mechanics_boundary(model, mat, data, mk), contacts, constants, keyword=mk)

    def dofs(name, g):
        return es.dofs_of([v for v in es.variables if v.name == name and v.domain is g])
    cols = [dofs(model.displacement_variable, mat)] + [dofs(model.contact_traction_variable, f) for f in fracs] \
        + [dofs(model.interface_displacement_variable, it) for it in intfs]
    return prob, np.concatenate(cols)


def _contact_constants(model, fracs):
    def scalar(op):
        return float(np.atleast_1d(_evaluated(model, op, 1))[0])
    return dict(numerical_constant=scalar(model.contact_mechanics_numerical_constant(fracs)),
                characteristic_traction=scalar(model.characteristic_contact_traction(fracs)),
                friction_coefficient=scalar(model.friction_coefficient(fracs)),
                dilation_angle=model.solid.dilation_angle, reference_gap=model.solid.fracture_gap,
                open_state_tolerance=model.numerical.open_state_tolerance)


def _fractured_problem(model, thermal: bool):
    """Shared part of the two fractured (thermo-)poromechanics bridges."""
    from .fractured_poromech import FractureCoupling
    mdg, es = model.mdg, model.equation_system
    mats, fracs = list(mdg.subdomains(dim=3)), list(mdg.subdomains(dim=2))
    if len(mats) != 1 or any(sd.dim < 2 for sd in mdg.subdomains()):
        raise NotImplementedError("one 3-D matrix subdomain and 2-D fractures without intersections are expected")
    mat = mats[0]
    fk, mk = model.darcy_keyword, model.stress_keyword
    kws = [fk, mk] + ([model.fourier_keyword] if thermal else [])
    data = _own_data(mdg.subdomain_data(mat), kws)
    a_res = model.solid.residual_aperture
    couplings, intfs, kappa_t = [], [], []
    for frac in fracs:
        intf = [it for it in mdg.interfaces() if mdg.interface_to_subdomain_pair(it)[1] is frac][0]
        fdata = _own_data(mdg.subdomain_data(frac), [fk] + ([model.fourier_keyword] if thermal else []))
        k_now = np.asarray(fdata[PARAMETERS][fk]["second_order_tensor"].values, float)
        a_now = _evaluated(model, model.specific_volume([frac]), frac.num_cells)       # the tensor holds k x specific volume
        proj = {name: getattr(intf, name)() for name in (
            "mortar_to_primary_avg", "primary_to_mortar_int", "mortar_to_secondary_avg", "secondary_to_mortar_int",
            "mortar_to_primary_int", "primary_to_mortar_avg", "mortar_to_secondary_int", "secondary_to_mortar_avg")}
        rot = mdg.subdomain_data(frac)["tangential_normal_projection"].project_tangential_normal(frac.num_cells)
        fbc = None
        if np.any(np.asarray(frac.tags["domain_boundary_faces"], bool)):      # the fracture reaches the domain boundary
            fl_ = _fluid(model, thermal)
            wf_, wef_ = _boundary_weights(model, frac, fl_, thermal)
            fft = model.bc_type_fluid_flux(frac)
            fbc = dict(flow=_face_values(model, frac, fdata[PARAMETERS][fk]["bc"], model.bc_values_pressure, model.bc_values_darcy_flux),
                       fluid_flux=_face_values(model, frac, fft, wf_, model.bc_values_fluid_flux), fluid_flux_type=fft)
            if thermal:
                eft = model.bc_type_enthalpy_flux(frac)
                fbc.update(fourier=_face_values(model, frac, fdata[PARAMETERS][model.fourier_keyword]["bc"],
                                                model.bc_values_temperature, model.bc_values_fourier_flux),
                           enthalpy_flux=_face_values(model, frac, eft, wef_, model.bc_values_enthalpy_flux),
                           enthalpy_flux_type=eft)
        couplings.append(FractureCoupling(frac, fdata, proj, sps.csr_matrix(intf.sign_of_mortar_sides(1)).diagonal(),
                                          intf.cell_volumes, rot, _evaluated(model, model.normal_permeability([intf]), intf.num_cells),
                                          k_now / a_now[None, None, :], bc=fbc))
        intfs.append(intf)
        if thermal:
            kappa_t.append(_evaluated(model, model.normal_thermal_conductivity([intf]), intf.num_cells))
    fluid = _fluid(model, thermal)
    so = model.solid
    solid = dict(reference_porosity=so.porosity, n_inv=_n_inv(model), residual_aperture=a_res)
    w, we = _boundary_weights(model, mat, fluid, thermal)
    ff = model.bc_type_fluid_flux(mat)
    prm = data[PARAMETERS]
    bc = dict(flow=_face_values(model, mat, prm[fk]["bc"], model.bc_values_pressure, model.bc_values_darcy_flux),
              mechanics=_mechanics_boundary(model, mat, data, mk),
              fluid_flux=_face_values(model, mat, ff, w, model.bc_values_fluid_flux), fluid_flux_type=ff)

    def dofs(name, g):
        return es.dofs_of([v for v in es.variables if v.name == name and v.domain is g])
    return mat, fracs, intfs, data, couplings, fluid, solid, bc, kappa_t, (w, we), dofs


def _row_map(model, layout_order):
    es = model.equation_system
    rows, r0 = {}, 0
    sizes = {}
    for eq, groups in layout_order:
        for g, mult in groups:
            sizes[(eq, id(g))] = mult * g.num_cells
    for eq in es.equations:
        for (name, gid), n in sizes.items():
            if name == eq:
                rows[(name, gid)] = np.arange(r0, r0 + n)
                r0 += n
    return np.concatenate([rows[(eq, id(g))] for eq, groups in layout_order for g, _ in groups])


def fractured_poromechanics_from_model(model):
    """``pp.Poromechanics`` on a fractured medium with frictional contact -> (``FracturedPoromechanics``, column_map,
    row_map)."""
    from .fractured_poromech import FracturedPoromechanics
    mat, fracs, intfs, data, couplings, fluid, solid, bc, _, _, dofs = _fractured_problem(model, False)
    prob = FracturedPoromechanics(mat, data, couplings, fluid, solid, _contact_constants(model, fracs), bc,
                                  flow_keyword=model.darcy_keyword, mechanics_keyword=model.stress_keyword)
    prob.mobility_keyword = "b200_mobility"
    cols = [dofs(model.pressure_variable, mat)] + [dofs(model.pressure_variable, f) for f in fracs] \
        + [dofs(model.displacement_variable, mat)] + [dofs(model.contact_traction_variable, f) for f in fracs] \
        + [dofs(model.interface_darcy_flux_variable, it) for it in intfs] \
        + [dofs(model.interface_displacement_variable, it) for it in intfs]
    order = [("mass_balance_equation", [(mat, 1)] + [(f, 1) for f in fracs]), ("momentum_balance_equation", [(mat, 3)]),
             ("interface_darcy_flux_equation", [(it, 1) for it in intfs]),
             ("interface_force_balance_equation", [(it, 3) for it in intfs]),
             ("normal_fracture_deformation_equation", [(f, 1) for f in fracs]),
             ("tangential_fracture_deformation_equation", [(f, 2) for f in fracs])]
    return prob, np.concatenate(cols), _row_map(model, order)


def fractured_thermoporomechanics_from_model(model):
    """``pp.Thermoporomechanics`` on a fractured medium with frictional contact (BASELINE config[4]) ->
    (``FracturedThermoporomechanics``, column_map, row_map)."""
    from .fractured_thm import FracturedThermoporomechanics
    mat, fracs, intfs, data, couplings, fluid, solid, bc, kappa_t, (w, we), dofs = _fractured_problem(model, True)
    so = model.solid
    solid.update(biot_coefficient=so.biot_coefficient, thermal_expansion=so.thermal_expansion,
                 heat_capacity=so.specific_heat_capacity, conductivity=so.thermal_conductivity, density=so.density)
    tk = model.fourier_keyword
    ef = model.bc_type_enthalpy_flux(mat)
    bc.update(fourier=_face_values(model, mat, data[PARAMETERS][tk]["bc"], model.bc_values_temperature, model.bc_values_fourier_flux),
              enthalpy_flux=_face_values(model, mat, ef, we, model.bc_values_enthalpy_flux), enthalpy_flux_type=ef)
    prob = FracturedThermoporomechanics(mat, data, couplings, fluid, solid, _contact_constants(model, fracs), bc, kappa_t,
                                        flow_keyword=model.darcy_keyword, fourier_keyword=tk,
                                        mechanics_keyword=model.stress_keyword, thermal_keyword=model.enthalpy_keyword)
    prob.mobility_keyword, prob.enthalpy_upwind_keyword = "b200_mobility", "b200_enthalpy_upwind"
    pv, tv = model.pressure_variable, model.temperature_variable
    cols = [dofs(pv, mat)] + [dofs(pv, f) for f in fracs] + [dofs(tv, mat)] + [dofs(tv, f) for f in fracs] \
        + [dofs(model.displacement_variable, mat)] + [dofs(model.contact_traction_variable, f) for f in fracs] \
        + [dofs(model.interface_darcy_flux_variable, it) for it in intfs] \
        + [dofs(model.interface_fourier_flux_variable, it) for it in intfs] \
        + [dofs(model.interface_enthalpy_flux_variable, it) for it in intfs] \
        + [dofs(model.interface_displacement_variable, it) for it in intfs]
    order = [("mass_balance_equation", [(mat, 1)] + [(f, 1) for f in fracs]),
             ("energy_balance_equation", [(mat, 1)] + [(f, 1) for f in fracs]), ("momentum_balance_equation", [(mat, 3)]),
             ("interface_darcy_flux_equation", [(it, 1) for it in intfs]),
             ("interface_fourier_flux_equation", [(it, 1) for it in intfs]),
             ("interface_enthalpy_flux_equation", [(it, 1) for it in intfs]),
             ("interface_force_balance_equation", [(it, 3) for it in intfs]),
             ("normal_fracture_deformation_equation", [(f, 1) for f in fracs]),
             ("tangential_fracture_deformation_equation", [(f, 2) for f in fracs])]
    return prob, np.concatenate(cols), _row_map(model, order)
