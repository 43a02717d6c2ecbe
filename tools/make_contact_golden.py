"""Generate tests/golden/contact_model.npz: frictional contact on a fracture, the unmodified reference's
``pp.MomentumBalance`` (MPSA elasticity in the matrix with the fracture faces as internal Dirichlet boundary, interface
displacements, contact traction; force balance on the interface; the semismooth normal / tangential complementarity laws of
models/contact_mechanics.py:80-245 with Coulomb friction and shear dilation) -- the contact part of BASELINE config[4] in
miniature: a compressed and sheared fracture in the sliding regime.  Stored: grids, parameters, the geometric pieces of the
interface (scalar mortar projections, side signs, local fracture coordinates), the Jacobian / right-hand side at the second
Newton iterate, the residual history and the converged state.   python tools/make_contact_golden.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_porepy  # noqa: E402
from make_golden import grid_arrays  # noqa: E402
from make_mdflow_golden import put_csr, rect  # noqa: E402

pp = load_porepy()
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class Model(pp.MomentumBalance):
    def set_domain(self):
        self._domain = pp.Domain({"xmin": 0, "xmax": 1, "ymin": 0, "ymax": 1, "zmin": 0, "zmax": 1})

    def grid_type(self):
        return "cartesian"

    def meshing_arguments(self):
        return {"cell_size": 0.25}

    def set_fractures(self):
        self._fractures = [pp.PlaneFracture(rect(0, 0.5, 0.25, 0.75))]

    def stiffness_tensor(self, sd):
        rng = np.random.default_rng(11 + sd.num_cells)
        return pp.FourthOrderTensor(1.5 * np.exp(0.2 * rng.standard_normal(sd.num_cells)),
                                    2.0 * np.exp(0.2 * rng.standard_normal(sd.num_cells)))

    def bc_type_mechanics(self, sd):
        s = self.domain_boundary_sides(sd)
        bc = pp.BoundaryConditionVectorial(sd, s.west + s.east, "dir")
        bc.internal_to_dirichlet(sd)
        return bc

    scenario = "sliding"

    def bc_values_displacement(self, bg):
        s = self.domain_boundary_sides(bg)
        v = np.zeros((3, bg.num_cells))
        if self.scenario == "sliding":
            v[0, s.east] = -0.01 * (1 + 0.3 * bg.cell_centers[2, s.east])     # compress across the fracture, unevenly
            v[1, s.east] = 0.02                                                # and shear it
            v[2, s.east] = 0.005 * bg.cell_centers[1, s.east]
        elif self.scenario == "sticking":                                      # strong compression, little shear
            v[0, s.east] = -0.02
            v[1, s.east] = 0.002 * bg.cell_centers[2, s.east]
        elif self.scenario == "open":                                          # pull the fracture open (+ some shear)
            v[0, s.east] = 0.01 * (1 + 0.5 * bg.cell_centers[1, s.east])
            v[2, s.east] = 0.004
        elif self.scenario == "mixed":                                         # a rotation-like load: part closes, part opens
            v[0, s.east] = 0.03 * (bg.cell_centers[2, s.east] - 0.5)
            v[1, s.east] = 0.01
        return v.ravel("F")


def main(scenario="sliding", name="contact_model"):
    solid = pp.SolidConstants(lame_lambda=2.0, shear_modulus=1.5, friction_coefficient=0.4, fracture_gap=1e-4,
                              dilation_angle=0.1)
    m = Model({"times_to_export": [], "time_manager": pp.TimeManager([0, 1.0], 1.0, constant_dt=True),
               "material_constants": {"solid": solid}})
    m.scenario = scenario
    m.prepare_simulation()
    es, mdg = m.equation_system, m.mdg
    mat, frac, intf = mdg.subdomains(dim=3)[0], mdg.subdomains(dim=2)[0], mdg.interfaces()[0]
    assert list(es.equations) == ["momentum_balance_equation", "interface_force_balance_equation",
                                  "normal_fracture_deformation_equation", "tangential_fracture_deformation_equation"]

    def dofs(name):
        return es.dofs_of([v for v in es.variables if v.name == name])
    d = {f"matrix__{k}": v for k, v in grid_arrays(mat).items()}
    d.update({f"fracture__{k}": v for k, v in grid_arrays(frac).items()})
    data = mdg.subdomain_data(mat)
    bcm = data[pp.PARAMETERS]["mechanics"]["bc"] if "bc" in data[pp.PARAMETERS].get("mechanics", {}) else m.bc_type_mechanics(mat)
    m.time_manager.increase_time()
    m.time_manager.increase_time_index()
    m.before_nonlinear_loop()
    x_prev = es.get_variable_values(time_step_index=0)
    norms = []
    for it in range(20):
        m.before_nonlinear_iteration()
        m.assemble_linear_system()
        A, b = m.linear_system
        norms.append(np.linalg.norm(b))
        if it == 1:
            d["iterate"] = es.get_variable_values(iterate_index=0)
            d["iterate_rhs"] = b.copy()
            put_csr(d, "iterate_jacobian", A)
        if norms[-1] < 1e-11 * norms[0]:
            break
        m.after_nonlinear_iteration(m.solve_linear_system())
    bcm = mdg.subdomain_data(mat)[pp.PARAMETERS]["mechanics"]["bc"]
    bg = mdg.subdomain_to_boundary_grid(mat)
    proj3 = sps.kron(bg.projection(), sps.eye(3)).tocsr()

    def scalar(op):
        v = es.evaluate(op)
        return float(np.atleast_1d(getattr(v, "val", v))[0])
    rot = mdg.subdomain_data(frac)["tangential_normal_projection"].project_tangential_normal(frac.num_cells)
    d.update(previous=x_prev, solution=es.get_variable_values(iterate_index=0), residual_norms=np.array(norms),
             column_map=np.concatenate([dofs("u"), dofs("contact_traction"), dofs("u_interface")]),
             C=mdg.subdomain_data(mat)[pp.PARAMETERS]["mechanics"]["fourth_order_tensor"].values,
             mech_is_dir=bcm.is_dir, mech_is_neu=bcm.is_neu, mech_is_rob=bcm.is_rob, mech_is_internal=bcm.is_internal,
             mech_bc_values=np.where(bcm.is_dir.ravel("F"), proj3.T @ m.bc_values_displacement(bg),
                                     proj3.T @ m.bc_values_stress(bg)),
             mortar_sign=sps.csr_matrix(intf.sign_of_mortar_sides(1)).diagonal(), mortar_volumes=intf.cell_volumes,
             numerical_constant=np.float64(scalar(m.contact_mechanics_numerical_constant([frac]))),
             characteristic_traction=np.float64(scalar(m.characteristic_contact_traction([frac]))),
             friction_coefficient=np.float64(scalar(m.friction_coefficient([frac]))),
             dilation_angle=np.float64(m.solid.dilation_angle), reference_gap=np.float64(m.solid.fracture_gap),
             open_state_tolerance=np.float64(m.numerical.open_state_tolerance))
    put_csr(d, "local_coordinates", rot)
    for key in ("mortar_to_primary_avg", "primary_to_mortar_int", "mortar_to_secondary_avg", "secondary_to_mortar_int"):
        put_csr(d, key, getattr(intf, key)())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    t = d["solution"][dofs("contact_traction")].reshape(-1, 3)
    print(name, "dofs", es.num_dofs(), "Newton residuals", ["%.2e" % v for v in norms])
    print("   contact traction (t1, t2, n) per fracture cell:\\n", t, "\\n   |t_t| / (mu |t_n|):",
          np.linalg.norm(t[:, :2], axis=1) / (0.4 * np.abs(t[:, 2])))


if __name__ == "__main__":
    main("sliding", "contact_model")
    main("sticking", "contact_sticking")
    main("open", "contact_open")
    main("mixed", "contact_mixed")
