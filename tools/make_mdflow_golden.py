"""Generate tests/golden/mdflow_*.npz: a whole mixed-dimensional single-phase flow problem of the unmodified reference
(run in the build container; the GPU box never sees /root/reference).

``pp.SinglePhaseFlow`` on a Cartesian matrix cut by three grid-aligned fractures (3-D matrix, three 2-D planes, six 1-D
intersection lines, one 0-D point, twelve + interfaces), anisotropic heterogeneous permeability, Dirichlet pressure on
two sides.  The fixture holds, per subdomain, the grid arrays and the parameters the reference handed to its flux
discretization (``data[pp.PARAMETERS]["flow"]``), per interface the mortar projections and the coefficients of the
interface law (reference models/constitutive_laws.py:1032-1076), and the reference's own global Jacobian, right-hand
side (``EquationSystem.assemble`` at the zero state, numerics/ad/equation_system.py:1579-1713) and converged solution.

    python tools/make_mdflow_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_porepy  # noqa: E402
from make_golden import grid_arrays  # noqa: E402

pp = load_porepy()
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def put_csr(d: dict, key: str, m) -> None:
    m = sps.csr_matrix(m)
    m.sum_duplicates()
    m.sort_indices()
    d[key + "__data"], d[key + "__indices"], d[key + "__indptr"] = m.data, m.indices.astype(np.int32), m.indptr.astype(np.int32)
    d[key + "__shape"] = np.array(m.shape, dtype=np.int64)


def rect(axis, at, lo, hi):
    """Corners of the axis-aligned rectangle {x_axis = at} x [lo, hi]^2."""
    o = [a for a in range(3) if a != axis]
    p = np.zeros((3, 4))
    p[axis] = at
    p[o[0]] = [lo, hi, hi, lo]
    p[o[1]] = [lo, lo, hi, hi]
    return p


def make_model(n, fracs, source, constants=None, dt=1.0, base=None, mixin=None):
    bases = ((mixin,) if mixin is not None else ()) + (base or pp.SinglePhaseFlow,)

    class Model(*bases):
        def set_domain(self):
            self._domain = pp.Domain({"xmin": 0, "xmax": 1, "ymin": 0, "ymax": 1, "zmin": 0, "zmax": 1})

        def grid_type(self):
            return "cartesian"

        def meshing_arguments(self):
            return {"cell_size": 1.0 / n}

        def set_fractures(self):
            self._fractures = [pp.PlaneFracture(f) for f in fracs]

        def permeability(self, subdomains):
            vals = []
            for sd in subdomains:
                rng = np.random.default_rng(1000 * sd.dim + sd.num_cells)
                nc = sd.num_cells
                t = np.zeros((3, 3, nc))
                scale = 1.0 if sd.dim == 3 else 50.0          # conductive fractures
                t[0, 0], t[1, 1], t[2, 2] = scale * (1 + rng.random((3, nc)))
                o = 0.3 * scale * rng.random((3, nc))
                t[0, 1] = t[1, 0] = o[0]
                t[0, 2] = t[2, 0] = o[1]
                t[1, 2] = t[2, 1] = o[2]
                vals.append(t.reshape(9, nc).ravel("F"))
            return pp.wrap_as_dense_ad_array(np.hstack(vals) if vals else np.zeros(0), name="permeability")

        def normal_permeability(self, interfaces):
            vals = [3.0 + np.cos(np.arange(i.num_cells)) for i in interfaces]
            return pp.wrap_as_dense_ad_array(np.hstack(vals) if vals else np.zeros(0), name="normal_permeability")

        def bc_type_darcy_flux(self, sd):
            sides = self.domain_boundary_sides(sd)
            return pp.BoundaryCondition(sd, sides.west + sides.east, "dir")

        def bc_values_pressure(self, bg):
            sides = self.domain_boundary_sides(bg)
            v = np.zeros(bg.num_cells)
            v[sides.west] = 1.0 + bg.cell_centers[1, sides.west]
            return v

        def bc_values_darcy_flux(self, bg):
            sides = self.domain_boundary_sides(bg)
            v = np.zeros(bg.num_cells)
            v[sides.top] = -0.1 * bg.cell_volumes[sides.top]       # inflow through the top
            return v

        def extra_source(self, sd):
            """Integrated cell sources on top of the interface inflow the reference adds itself
            (models/fluid_mass_balance.py ``fluid_source``)."""
            return source * sd.cell_volumes * np.sin(3 * sd.cell_centers[0]) if sd.dim == 3 else np.zeros(sd.num_cells)

        def fluid_source(self, subdomains):
            vals = [self.extra_source(sd) for sd in subdomains]
            ext = pp.wrap_as_dense_ad_array(np.hstack(vals) if vals else np.zeros(0), name="extra_source")
            return super().fluid_source(subdomains) + ext

        def bc_type_fluid_flux(self, sd):
            if constants is None:               # the linear fixtures keep the model's default
                return super().bc_type_fluid_flux(sd)
            sides = self.domain_boundary_sides(sd)
            return pp.BoundaryCondition(sd, sides.west + sides.east, "dir")

    params = {"times_to_export": [], "time_manager": pp.TimeManager([0, 1.0], dt, constant_dt=True)}
    if constants is not None:
        params["material_constants"] = constants
    return Model(params)


def full_bc_values(model, sd):
    """Face-wise boundary data as the flux discretization consumes it: pressures on Dirichlet faces, fluxes elsewhere."""
    mdg = model.mdg
    v = np.zeros(sd.num_faces)
    bg = mdg.subdomain_to_boundary_grid(sd)
    if bg is None or bg.num_cells == 0:
        return v
    bc = mdg.subdomain_data(sd)[pp.PARAMETERS]["flow"]["bc"]
    proj = bg.projection()
    return np.where(bc.is_dir, proj.T @ model.bc_values_pressure(bg), proj.T @ model.bc_values_darcy_flux(bg))


def scalar_field(model, op, n):
    v = model.equation_system.evaluate(op)
    return np.full(n, float(v)) if np.ndim(v) == 0 else np.asarray(v, float)


def collect(model) -> dict:
    """Grids, parameters, projections and interface coefficients of a prepared model (everything but results)."""
    es, mdg = model.equation_system, model.mdg
    sds, intfs = mdg.subdomains(), mdg.interfaces()
    d = {"num_subdomains": np.int64(len(sds)), "num_interfaces": np.int64(len(intfs))}
    off = 0
    for i, sd in enumerate(sds):
        data = mdg.subdomain_data(sd)
        prm = data[pp.PARAMETERS]["flow"]
        assert np.array_equal(es.dofs_of([v for v in es.variables if v.name == "pressure" and v.domain is sd]),
                              off + np.arange(sd.num_cells))
        off += sd.num_cells
        g = grid_arrays(sd) if sd.dim > 0 else dict(
            dim=np.int64(0), name=np.array(str(sd.name)), nodes=sd.nodes, cell_centers=sd.cell_centers,
            cell_volumes=sd.cell_volumes)
        for k, v in g.items():
            d[f"sd{i}__{k}"] = v
        d[f"sd{i}__source"] = np.asarray(model.extra_source(sd), float)
        if sd.dim > 0:
            bc = prm["bc"]
            d[f"sd{i}__K"] = prm["second_order_tensor"].values
            for k in ("is_dir", "is_neu", "is_rob", "is_internal"):
                d[f"sd{i}__bc_{k}"] = getattr(bc, k)
            d[f"sd{i}__bc_robin_weight"] = np.asarray(bc.robin_weight, float)
            d[f"sd{i}__bc_values"] = full_bc_values(model, sd)
            d[f"sd{i}__tip_faces"] = np.asarray(sd.tags["tip_faces"], bool)
            d[f"sd{i}__domain_boundary_faces"] = np.asarray(sd.tags["domain_boundary_faces"], bool)
            d[f"sd{i}__ambient_dimension"] = np.int64(prm.get("ambient_dimension", 3))
    index = {sd: i for i, sd in enumerate(sds)}
    for j, it in enumerate(intfs):
        h, l = mdg.interface_to_subdomain_pair(it)
        assert np.array_equal(es.dofs_of([v for v in es.variables if v.name == "interface_darcy_flux" and v.domain is it]),
                              off + np.arange(it.num_cells))
        off += it.num_cells
        d[f"if{j}__primary"], d[f"if{j}__secondary"] = np.int64(index[h]), np.int64(index[l])
        put_csr(d, f"if{j}__mortar_to_primary_int", it.mortar_to_primary_int())
        put_csr(d, f"if{j}__primary_to_mortar_avg", it.primary_to_mortar_avg())
        put_csr(d, f"if{j}__mortar_to_secondary_int", it.mortar_to_secondary_int())
        put_csr(d, f"if{j}__secondary_to_mortar_avg", it.secondary_to_mortar_avg())
        d[f"if{j}__normal_permeability"] = scalar_field(model, model.normal_permeability([it]), it.num_cells)
        d[f"if{j}__cell_volumes"] = it.cell_volumes * scalar_field(model, model.specific_volume([it]), it.num_cells)
        d[f"if{j}__secondary_aperture"] = scalar_field(model, model.aperture([l]), l.num_cells)
    assert off == es.num_dofs()
    return d


def export(name, n, fracs, source=0.7):
    model = make_model(n, fracs, source)
    model.prepare_simulation()
    es, mdg = model.equation_system, model.mdg
    A0, b0 = es.assemble()
    d = collect(model)
    pp.run_time_dependent_model(model, {"prepare_simulation": False})
    A1, b1 = es.assemble()
    assert abs(A1 - A0).max() == 0.0 and np.linalg.norm(b1) < 1e-10 * np.linalg.norm(b0)
    x = es.get_variable_values(iterate_index=0)
    sds, intfs = mdg.subdomains(), mdg.interfaces()
    d.update(solution=x, rhs=b0)
    put_csr(d, "jacobian", A0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "dofs", A0.shape[0], "nnz", A0.nnz, "subdomains", [(s.dim, s.num_cells) for s in sds],
          "interfaces", [(i.dim, i.num_cells) for i in intfs], "|b|", np.linalg.norm(b0), "ptp(x)", np.ptp(x))


def export_nonlinear(name, n, fracs, source=0.3):
    """Compressible fluid (density rho0 exp(c (p - p_ref)), upwinded mobility rho / mu, storage term): one implicit time
    step of the reference's Newton loop, driven by hand (models/solution_strategy.py: ``before_nonlinear_iteration`` ->
    ``assemble_linear_system`` -> ``solve_linear_system`` -> ``after_nonlinear_iteration``).  Stored: the state, Jacobian
    and right-hand side in front of the third linear solve (upwind directions from that same iterate), the converged
    state, and what the nonlinear terms need beyond the linear fixture."""
    fluid = pp.FluidComponent(compressibility=0.05, viscosity=1.3, density=1.7)
    solid = pp.SolidConstants(porosity=0.2, residual_aperture=0.05)
    model = make_model(n, fracs, source, constants={"fluid": fluid, "solid": solid}, dt=0.25)
    model.prepare_simulation()
    es, mdg = model.equation_system, model.mdg
    d = collect(model)
    model.time_manager.increase_time()
    model.time_manager.increase_time_index()
    model.before_nonlinear_loop()
    x_prev = es.get_variable_values(time_step_index=0)
    norms = []
    for it in range(12):
        model.before_nonlinear_iteration()
        model.assemble_linear_system()
        A, b = model.linear_system
        norms.append(np.linalg.norm(b))
        if it == 2:
            d["iterate"] = es.get_variable_values(iterate_index=0)
            d["iterate_rhs"] = b.copy()
            put_csr(d, "iterate_jacobian", A)
        if norms[-1] < 1e-13 * norms[0]:
            break
        model.after_nonlinear_iteration(model.solve_linear_system())
    fl = model.fluid.reference_component
    d.update(previous=x_prev, solution=es.get_variable_values(iterate_index=0), residual_norms=np.array(norms),
             dt=np.float64(model.time_manager.dt), compressibility=np.float64(fl.compressibility),
             density=np.float64(fl.density), viscosity=np.float64(fl.viscosity),
             reference_pressure=np.float64(model.reference_variable_values.pressure))
    for i, sd in enumerate(mdg.subdomains()):
        sv = scalar_field(model, model.specific_volume([sd]), sd.num_cells)
        phi = scalar_field(model, model.porosity([sd]), sd.num_cells)
        d[f"sd{i}__storage"] = sd.cell_volumes * sv * phi
        if sd.dim == 0:
            continue
        bc = model.bc_type_fluid_flux(sd)
        d[f"sd{i}__ff_is_dir"], d[f"sd{i}__ff_is_neu"] = bc.is_dir, bc.is_neu
        w = np.zeros(sd.num_faces)
        bg = mdg.subdomain_to_boundary_grid(sd)
        if bg is not None and bg.num_cells > 0:
            proj = bg.projection()
            pb_ = proj.T @ model.bc_values_pressure(bg)
            rho = fl.density * np.exp(fl.compressibility * (pb_ - model.reference_variable_values.pressure))
            w = np.where(bc.is_dir, rho / fl.viscosity, proj.T @ model.bc_values_fluid_flux(bg))
        d[f"sd{i}__ff_values"] = w
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "dofs", es.num_dofs(), "Newton residuals", ["%.2e" % v for v in norms])


def export_thermal(name, n, fracs, source=0.3):
    """``pp.MassAndEnergyBalance`` on the network (compressible, thermally expanding fluid; Fourier and upwinded enthalpy
    fluxes on subdomains and interfaces): the fixture of ``export_nonlinear`` plus the thermal parameters, and the index
    arrays that map the reference's dof / equation numbering (interleaved per grid) to [p | T | lambda | eta | eps] and
    [mass | energy | Darcy law | Fourier law | enthalpy law]."""
    fluid = pp.FluidComponent(compressibility=0.05, viscosity=1.3, density=1.7, thermal_expansion=0.03,
                              specific_heat_capacity=2.0, thermal_conductivity=0.7)
    solid = pp.SolidConstants(porosity=0.2, residual_aperture=0.05, thermal_expansion=0.02, specific_heat_capacity=1.5,
                              thermal_conductivity=1.1, density=2.5, normal_permeability=2.0)

    class Thermal:
        def bc_type_fourier_flux(self, sd):
            s = self.domain_boundary_sides(sd)
            return pp.BoundaryCondition(sd, s.west + s.east, "dir")

        def bc_type_enthalpy_flux(self, sd):
            s = self.domain_boundary_sides(sd)
            return pp.BoundaryCondition(sd, s.west + s.east, "dir")

        def bc_values_temperature(self, bg):
            s = self.domain_boundary_sides(bg)
            v = np.zeros(bg.num_cells)
            v[s.west] = 0.5 + 0.2 * bg.cell_centers[2, s.west]
            return v
    model = make_model(n, fracs, source, constants={"fluid": fluid, "solid": solid}, dt=0.25,
                       base=pp.MassAndEnergyBalance, mixin=Thermal)
    model.prepare_simulation()
    es, mdg = model.equation_system, model.mdg
    sds, intfs = mdg.subdomains(), mdg.interfaces()
    # grids, flow parameters, projections: as for the flow fixtures (``collect`` asserts the flow model's dof layout)
    d = {"num_subdomains": np.int64(len(sds)), "num_interfaces": np.int64(len(intfs))}
    index = {sd: i for i, sd in enumerate(sds)}
    fl = model.fluid.reference_component
    p_ref, t_ref = model.reference_variable_values.pressure, model.reference_variable_values.temperature

    def dofs(name, g):
        return es.dofs_of([v for v in es.variables if v.name == name and v.domain is g])
    cols, rows, r0 = [], {}, 0
    for eq in es.equations:
        for g in (sds if eq in ("mass_balance_equation", "energy_balance_equation") else
                  intfs if eq.startswith("interface") else []):
            rows[(eq, g)] = np.arange(r0, r0 + g.num_cells)
            r0 += g.num_cells
    for var, grids in (("pressure", sds), ("temperature", sds), ("interface_darcy_flux", intfs),
                       ("interface_fourier_flux", intfs), ("interface_enthalpy_flux", intfs)):
        cols += [dofs(var, g) for g in grids]
    order = [("mass_balance_equation", sds), ("energy_balance_equation", sds), ("interface_darcy_flux_equation", intfs),
             ("interface_fourier_flux_equation", intfs), ("interface_enthalpy_flux_equation", intfs)]
    d["column_map"] = np.concatenate(cols)          # my unknown k is the reference's dof column_map[k]
    d["row_map"] = np.concatenate([rows[(eq, g)] for eq, grids in order for g in grids])
    for i, sd in enumerate(sds):
        data = mdg.subdomain_data(sd)
        prm = data[pp.PARAMETERS]
        g = grid_arrays(sd) if sd.dim > 0 else dict(dim=np.int64(0), name=np.array(str(sd.name)), nodes=sd.nodes,
                                                     cell_centers=sd.cell_centers, cell_volumes=sd.cell_volumes)
        for k, v in g.items():
            d[f"sd{i}__{k}"] = v
        d[f"sd{i}__source"] = np.asarray(model.extra_source(sd), float)
        d[f"sd{i}__volume"] = sd.cell_volumes * scalar_field(model, model.specific_volume([sd]), sd.num_cells)
        d[f"sd{i}__porosity"] = scalar_field(model, model.porosity([sd]), sd.num_cells)
        if sd.dim == 0:
            continue
        bg = mdg.subdomain_to_boundary_grid(sd)
        has = bg is not None and bg.num_cells > 0
        proj = bg.projection() if has else None
        pb_ = proj.T @ model.bc_values_pressure(bg) if has else np.zeros(sd.num_faces)
        tb = proj.T @ model.bc_values_temperature(bg) if has else np.zeros(sd.num_faces)
        rho_b = fl.density * np.exp(fl.compressibility * (pb_ - p_ref) - fl.thermal_expansion * (tb - t_ref))

        def comb(bc, dirv, neu):
            return np.where(bc.is_dir, dirv, proj.T @ neu(bg)) if has else np.zeros(sd.num_faces)
        for key, kw in (("flow", "flow"), ("fourier", "fourier_discretization")):
            bc = prm[kw]["bc"]
            d[f"sd{i}__{key}_K"] = prm[kw]["second_order_tensor"].values
            for f in ("is_dir", "is_neu", "is_rob", "is_internal"):
                d[f"sd{i}__{key}_bc_{f}"] = getattr(bc, f)
            d[f"sd{i}__{key}_bc_robin_weight"] = np.asarray(bc.robin_weight, float)
        d[f"sd{i}__flow_bc_values"] = comb(prm["flow"]["bc"], pb_, model.bc_values_darcy_flux)
        d[f"sd{i}__fourier_bc_values"] = comb(prm["fourier_discretization"]["bc"], tb, model.bc_values_fourier_flux)
        for key, bc, dirv, neu in (("ff", model.bc_type_fluid_flux(sd), rho_b / fl.viscosity, model.bc_values_fluid_flux),
                                   ("ef", model.bc_type_enthalpy_flux(sd),
                                    fl.specific_heat_capacity * (tb - t_ref) * rho_b / fl.viscosity,
                                    model.bc_values_enthalpy_flux)):
            d[f"sd{i}__{key}_is_dir"], d[f"sd{i}__{key}_is_neu"] = bc.is_dir, bc.is_neu
            d[f"sd{i}__{key}_values"] = comb(bc, dirv, neu)
        d[f"sd{i}__tip_faces"] = np.asarray(sd.tags["tip_faces"], bool)
        d[f"sd{i}__domain_boundary_faces"] = np.asarray(sd.tags["domain_boundary_faces"], bool)
        d[f"sd{i}__ambient_dimension"] = np.int64(prm["flow"].get("ambient_dimension", 3))
    for j, it in enumerate(intfs):
        h, l = mdg.interface_to_subdomain_pair(it)
        d[f"if{j}__primary"], d[f"if{j}__secondary"] = np.int64(index[h]), np.int64(index[l])
        put_csr(d, f"if{j}__mortar_to_primary_int", it.mortar_to_primary_int())
        put_csr(d, f"if{j}__primary_to_mortar_avg", it.primary_to_mortar_avg())
        put_csr(d, f"if{j}__mortar_to_secondary_int", it.mortar_to_secondary_int())
        put_csr(d, f"if{j}__secondary_to_mortar_avg", it.secondary_to_mortar_avg())
        d[f"if{j}__normal_permeability"] = scalar_field(model, model.normal_permeability([it]), it.num_cells)
        d[f"if{j}__normal_thermal_conductivity"] = scalar_field(model, model.normal_thermal_conductivity([it]), it.num_cells)
        d[f"if{j}__cell_volumes"] = it.cell_volumes * scalar_field(model, model.specific_volume([it]), it.num_cells)
        d[f"if{j}__secondary_aperture"] = scalar_field(model, model.aperture([l]), l.num_cells)
    model.time_manager.increase_time()
    model.time_manager.increase_time_index()
    model.before_nonlinear_loop()
    x_prev = es.get_variable_values(time_step_index=0)
    norms = []
    for it in range(15):
        model.before_nonlinear_iteration()
        model.assemble_linear_system()
        A, b = model.linear_system
        norms.append(np.linalg.norm(b))
        if it == 2:
            d["iterate"] = es.get_variable_values(iterate_index=0)
            d["iterate_rhs"] = b.copy()
            put_csr(d, "iterate_jacobian", A)
        if norms[-1] < 1e-12 * norms[0]:
            break
        model.after_nonlinear_iteration(model.solve_linear_system())
    so = model.solid
    d.update(previous=x_prev, solution=es.get_variable_values(iterate_index=0), residual_norms=np.array(norms),
             dt=np.float64(model.time_manager.dt), compressibility=np.float64(fl.compressibility),
             density=np.float64(fl.density), viscosity=np.float64(fl.viscosity),
             fluid_thermal_expansion=np.float64(fl.thermal_expansion), fluid_heat_capacity=np.float64(fl.specific_heat_capacity),
             reference_pressure=np.float64(p_ref), reference_temperature=np.float64(t_ref),
             solid_heat_capacity=np.float64(so.specific_heat_capacity), solid_density=np.float64(so.density))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "dofs", es.num_dofs(), "Newton residuals", ["%.2e" % v for v in norms])


if __name__ == "__main__":
    export("mdflow_three_fractures", 6, [rect(0, 0.5, 1 / 6, 5 / 6), rect(1, 0.5, 1 / 6, 5 / 6), rect(2, 0.5, 1 / 6, 5 / 6)])
    export("mdflow_one_fracture", 4, [rect(0, 0.5, 0.25, 0.75)], source=0.0)
    export_nonlinear("mdflownl_three_fractures", 6, [rect(0, 0.5, 1 / 6, 5 / 6), rect(1, 0.5, 1 / 6, 5 / 6), rect(2, 0.5, 1 / 6, 5 / 6)])
    export_nonlinear("mdflownl_one_fracture", 4, [rect(0, 0.5, 0.25, 0.75)])
    export_thermal("mdthermal_three_fractures", 6, [rect(0, 0.5, 1 / 6, 5 / 6), rect(1, 0.5, 1 / 6, 5 / 6), rect(2, 0.5, 1 / 6, 5 / 6)])
    export_thermal("mdthermal_one_fracture", 4, [rect(0, 0.5, 0.25, 0.75)])
