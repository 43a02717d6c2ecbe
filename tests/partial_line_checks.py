"""Checks shared by the CPU (host build) and GPU legs: the reference's in-place partial update and the 1-D line."""
import numpy as np
import scipy.sparse as sps

import porepy_b200 as pb
from cases import load_case, max_rel_err


def check_partial_update(name, tol=1e-10):
    """Old parameters -> full pass; then the changed tensors with ``specified_cells`` + ``update_discretization``:
    the stored matrices must equal what the reference's ``update_discretization`` (partial_update_discretization,
    _fvutils.py:1090-1257) ended up with."""
    c = load_case(name)
    cells = c.raw["modified_cells"]
    if c.kind == "partial_mpfa":
        kw, discr = "flow", pb.Mpfa("flow")
        old = {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]), "bc": c.bc, "mpfa_eta": c.eta}
        new = dict(old, second_order_tensor=pb.SecondOrderTensor.from_values(c.raw["K2"]))
    else:
        kw, discr = "mech", pb.Mpsa("mech")
        old = {"fourth_order_tensor": pb.FourthOrderTensor.from_values(c.raw["C"]), "bc": c.bc, "mpsa_eta": c.eta}
        new = dict(old, fourth_order_tensor=pb.FourthOrderTensor.from_values(c.raw["C2"]))
    data = pb.initialize_data({}, kw, old)
    discr.discretize(c.g, data)
    stored = dict(data[pb.DISCRETIZATION_MATRICES][kw])
    upd = pb.initialize_data({}, kw, dict(new, specified_cells=cells, update_discretization=True))
    upd[pb.DISCRETIZATION_MATRICES][kw] = stored
    discr.discretize(c.g, upd)
    err, key = max_rel_err(c.mats, upd[pb.DISCRETIZATION_MATRICES][kw])
    assert err < tol, (name, key, err)
    return err


def check_line(name="line1d_tilted", tol=1e-13):
    """1-D grid on a tilted line: ``pb.Mpfa`` (TPFA delegation, 3 ambient components), ``pb.Mpsa`` and ``pb.Upwind``
    against the reference's matrices."""
    c = load_case(name)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]),
                                           "bc": c.bc, "ambient_dimension": 3})
    pb.Mpfa("flow").discretize(c.g, data)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("tpfa_")}
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < tol, (key, err)
    md = pb.initialize_data({}, "mech", {"fourth_order_tensor": pb.FourthOrderTensor(c.raw["mu"], c.raw["lmbda"]),
                                         "bc": pb.BoundaryConditionVectorial(c.g)})
    pb.Mpsa("mech").discretize(c.g, md)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("mpsa_")}
    err2, key = max_rel_err(ref, md[pb.DISCRETIZATION_MATRICES]["mech"])
    assert err2 < tol, (key, err2)
    td = pb.initialize_data({}, "transport", {"bc": c.bc, "darcy_flux": c.raw["darcy_flux"]})
    pb.Upwind("transport").discretize(c.g, td)
    M = td[pb.DISCRETIZATION_MATRICES]["transport"]
    for ref_key, key in (("upwind", "transport"), ("bound_transport_dir", "rhs_dir"), ("bound_transport_neu", "rhs_neu")):
        assert abs(sps.csr_matrix(c.mats[ref_key]) - M[key]).sum() == 0, key
    return max(err, err2)
