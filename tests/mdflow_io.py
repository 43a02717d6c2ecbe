"""Load the mixed-dimensional flow fixtures of tools/make_mdflow_golden.py into ``porepy_b200.mdflow`` records."""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200.grid import Grid
from porepy_b200.mdflow import MdInterface, MdSubdomain, MixedDimensionalFlow
from golden_io import GOLDEN_DIR


def _csr(d, key):
    return sps.csr_matrix((d[key + "__data"], d[key + "__indices"], d[key + "__indptr"]), shape=tuple(d[key + "__shape"]))


def load_mdflow(name: str):
    """(MixedDimensionalFlow, reference Jacobian, reference rhs, reference solution)."""
    return _load_linear_parts(name)


def _load_linear_parts(name: str):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    subs = []
    for i in range(int(d["num_subdomains"])):
        sub = {k[len(f"sd{i}__"):]: v for k, v in d.items() if k.startswith(f"sd{i}__")}
        if int(sub["dim"]) == 0:
            nc = sub["cell_volumes"].size
            g = SimpleNamespace(dim=0, num_cells=nc, num_faces=0, num_nodes=sub["nodes"].shape[1], nodes=sub["nodes"],
                                cell_centers=sub["cell_centers"], cell_volumes=sub["cell_volumes"],
                                cell_faces=sps.csc_matrix((0, nc)), name=str(sub["name"]))
            subs.append(MdSubdomain(g, {}, None, sub["source"]))
            continue
        g = Grid.from_arrays(sub)
        g.tags["tip_faces"] = np.asarray(sub["tip_faces"], bool)
        g.tags["domain_boundary_faces"] = np.asarray(sub["domain_boundary_faces"], bool)
        bc = SimpleNamespace(is_dir=sub["bc_is_dir"], is_neu=sub["bc_is_neu"], is_rob=sub["bc_is_rob"],
                             is_internal=sub["bc_is_internal"], robin_weight=sub["bc_robin_weight"], bc_type="scalar",
                             num_faces=g.num_faces)
        data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(sub["K"]),
                                               "bc": bc, "ambient_dimension": int(sub["ambient_dimension"])})
        subs.append(MdSubdomain(g, data, sub["bc_values"], sub["source"]))
    intfs = []
    for j in range(int(d["num_interfaces"])):
        p = f"if{j}__"
        intfs.append(MdInterface(int(d[p + "primary"]), int(d[p + "secondary"]), _csr(d, p + "mortar_to_primary_int"),
                                 _csr(d, p + "primary_to_mortar_avg"), _csr(d, p + "mortar_to_secondary_int"),
                                 _csr(d, p + "secondary_to_mortar_avg"), d[p + "normal_permeability"],
                                 d[p + "cell_volumes"], d[p + "secondary_aperture"]))
    jac = _csr(d, "jacobian") if "jacobian__data" in d else None
    return MixedDimensionalFlow(subs, intfs, "flow"), jac, d.get("rhs"), d["solution"]


def load_mdflow_nonlinear(name: str):
    """(CompressibleMixedDimensionalFlow, raw fixture) of tools/make_mdflow_golden.py ``export_nonlinear``."""
    from porepy_b200.mdflow_nl import CompressibleMixedDimensionalFlow
    lin, _, _, _ = _load_linear_parts(name)
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    storage, bcs, weights = [], [], []
    for i, s in enumerate(lin.subdomains):
        storage.append(d[f"sd{i}__storage"])
        if s.sd.num_faces == 0:
            bcs.append(None)
            weights.append(None)
            continue
        nf = s.sd.num_faces
        bcs.append(SimpleNamespace(is_dir=d[f"sd{i}__ff_is_dir"], is_neu=d[f"sd{i}__ff_is_neu"], is_rob=np.zeros(nf, bool),
                                   is_internal=np.asarray(s.sd.tags["fracture_faces"], bool), robin_weight=np.ones(nf),
                                   bc_type="scalar", num_faces=nf))
        weights.append(d[f"sd{i}__ff_values"])
    fluid = {k: float(d[k]) for k in ("compressibility", "density", "viscosity", "reference_pressure")}
    return CompressibleMixedDimensionalFlow(lin.subdomains, lin.interfaces, fluid, storage, bcs, weights), d


def load_mdthermal(name: str):
    """(MixedDimensionalMassEnergy, raw fixture) of tools/make_mdflow_golden.py ``export_thermal``."""
    from porepy_b200.mdthermal import MixedDimensionalMassEnergy
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    subs, volume, porosity, bcv, bct, sources = [], [], [], [], [], []

    def scalar_bc(sub, prefix, nf, internal):
        return SimpleNamespace(is_dir=sub[prefix + "is_dir"], is_neu=sub[prefix + "is_neu"],
                               is_rob=sub.get(prefix + "is_rob", np.zeros(nf, bool)),
                               is_internal=sub.get(prefix + "is_internal", internal),
                               robin_weight=sub.get(prefix + "robin_weight", np.ones(nf)), bc_type="scalar", num_faces=nf)
    for i in range(int(d["num_subdomains"])):
        sub = {k[len(f"sd{i}__"):]: v for k, v in d.items() if k.startswith(f"sd{i}__")}
        volume.append(sub["volume"])
        porosity.append(sub["porosity"])
        sources.append(sub["source"])
        if int(sub["dim"]) == 0:
            nc = sub["cell_volumes"].size
            g = SimpleNamespace(dim=0, num_cells=nc, num_faces=0, num_nodes=sub["nodes"].shape[1], nodes=sub["nodes"],
                                cell_centers=sub["cell_centers"], cell_volumes=sub["cell_volumes"],
                                cell_faces=sps.csc_matrix((0, nc)), name=str(sub["name"]))
            subs.append(MdSubdomain(g, {}))
            bcv.append(None)
            bct.append(None)
            continue
        g = Grid.from_arrays(sub)
        g.tags["tip_faces"] = np.asarray(sub["tip_faces"], bool)
        g.tags["domain_boundary_faces"] = np.asarray(sub["domain_boundary_faces"], bool)
        nf = g.num_faces
        internal = np.asarray(g.tags["fracture_faces"], bool)
        amb = int(sub["ambient_dimension"])
        data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(sub["flow_K"]),
                                               "bc": scalar_bc(sub, "flow_bc_", nf, internal), "ambient_dimension": amb})
        pb.initialize_data(data, "fourier", {"second_order_tensor": pb.SecondOrderTensor.from_values(sub["fourier_K"]),
                                             "bc": scalar_bc(sub, "fourier_bc_", nf, internal), "ambient_dimension": amb})
        subs.append(MdSubdomain(g, data))
        bcv.append(dict(flow=sub["flow_bc_values"], fourier=sub["fourier_bc_values"], fluid_flux=sub["ff_values"],
                        enthalpy_flux=sub["ef_values"]))
        bct.append(dict(fluid_flux=scalar_bc(sub, "ff_", nf, internal), enthalpy_flux=scalar_bc(sub, "ef_", nf, internal)))
    intfs, kappa_t = [], []
    for j in range(int(d["num_interfaces"])):
        p = f"if{j}__"
        intfs.append(MdInterface(int(d[p + "primary"]), int(d[p + "secondary"]), _csr(d, p + "mortar_to_primary_int"),
                                 _csr(d, p + "primary_to_mortar_avg"), _csr(d, p + "mortar_to_secondary_int"),
                                 _csr(d, p + "secondary_to_mortar_avg"), d[p + "normal_permeability"],
                                 d[p + "cell_volumes"], d[p + "secondary_aperture"]))
        kappa_t.append(d[p + "normal_thermal_conductivity"])
    fluid = dict(compressibility=d["compressibility"], density=d["density"], viscosity=d["viscosity"],
                 thermal_expansion=d["fluid_thermal_expansion"], heat_capacity=d["fluid_heat_capacity"],
                 reference_pressure=d["reference_pressure"], reference_temperature=d["reference_temperature"])
    solid = dict(density=d["solid_density"], heat_capacity=d["solid_heat_capacity"])
    prob = MixedDimensionalMassEnergy(subs, intfs, fluid, solid, volume, porosity, bcv, bct, kappa_t, sources)
    return prob, d
