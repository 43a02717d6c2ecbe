"""GPU leg of tests/test_tpfa_upwind.py: ``pb.Tpfa`` / ``pb.Upwind`` through the real per-face kernels
(``pb_facegrid`` + ``pb_tpfa`` / ``pb_upwind``, one thread per face) against golden outputs of the reference,
incl. a 1-D grid on a tilted line (the TPFA delegation of MPFA / MPSA)."""
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from cases import load_case, max_rel_err
from golden_io import case_names
from partial_line_checks import check_line

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", case_names("next_"))
def test_tpfa_and_upwind_on_the_device(name):
    c = load_case(name)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]),
                                           "bc": c.bc})
    pb.Tpfa("flow").discretize(c.g, data)
    ref = {k[5:]: v for k, v in c.mats.items() if k.startswith("tpfa_")}
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < 1e-13, (key, err)
    td = pb.initialize_data({}, "transport", {"bc": c.bc, "darcy_flux": c.raw["darcy_flux"]})
    pb.Upwind("transport").discretize(c.g, td)
    M = td[pb.DISCRETIZATION_MATRICES]["transport"]
    for ref_key, key in (("upwind", "transport"), ("bound_transport_dir", "rhs_dir"), ("bound_transport_neu", "rhs_neu")):
        assert abs(sps.csr_matrix(c.mats[ref_key]) - M[key]).sum() == 0, key


def test_line_grid_on_the_device():
    check_line()


def test_upwind_coupling_on_the_device():
    """``pb.UpwindCoupling`` (upwind.py:377-680): the interface masks from the device kernel, the trace operators and
    the assembled 3 x 3 block contribution -- against the reference's own class on a real mixed-dimensional grid when
    ``oracle/_ref`` is on the box, against the reference's formulas otherwise."""
    import numpy as np
    from types import SimpleNamespace
    g = pb.cart_grid_3d([4, 3, 3])
    rng = np.random.default_rng(4)
    n = 11
    lam = rng.standard_normal(n)
    lam[[2, 7]] = 0.0
    low = SimpleNamespace(dim=2)
    intf = SimpleNamespace(num_cells=n)
    di = pb.initialize_data({}, "transport", {"darcy_flux": lam})
    up = pb.UpwindCoupling("transport")
    up.discretize(g, low, intf, {}, {}, di)
    M = di[pb.DISCRETIZATION_MATRICES]["transport"]
    sgn = np.sign(lam)
    assert np.array_equal(M["flux"].diagonal(), sgn)
    assert np.array_equal(M["upwind_primary"].diagonal(), (sgn > 0).astype(float))
    assert np.array_equal(M["upwind_secondary"].diagonal(), 1 - (sgn > 0).astype(float))
    assert abs(M["inv_trace"] - abs(g.divergence(1))).sum() == 0 and abs(M["trace"] - abs(g.divergence(1)).T).sum() == 0
    assert abs(M["mortar_discr"] - sps.eye(n)).sum() == 0
    try:
        from oracle.ref_loader import load_porepy, reference_available
        if not reference_available():
            return
        pp = load_porepy()
    except Exception:
        return
    # the reference's class on a real interface (3-D matrix, one fracture plane)
    frac = pp.PlaneFracture(np.array([[0.5, 0.5, 0.5, 0.5], [0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0]]))
    mdg = pp.create_mdg("cartesian", {"cell_size": 0.25}, pp.create_fracture_network([frac], pp.domains.unit_cube_domain(3)))
    (intf_r, d_intf), = [(i, d) for i, d in mdg.interfaces(return_data=True) if i.dim == 2]
    sd_h, sd_l = mdg.interface_to_subdomain_pair(intf_r)
    lam = np.random.default_rng(5).standard_normal(intf_r.num_cells)
    ref_d = pp.initialize_data({}, "transport", {"darcy_flux": lam})
    mine_d = pp.initialize_data({}, "transport", {"darcy_flux": lam})
    ref, mine = pp.UpwindCoupling("transport"), pb.UpwindCoupling("transport")
    ref.discretize(sd_h, sd_l, intf_r, {}, {}, ref_d)
    mine.discretize(sd_h, sd_l, intf_r, {}, {}, mine_d)
    for key, m in ref_d[pp.DISCRETIZATION_MATRICES]["transport"].items():
        assert abs(sps.csr_matrix(m) - sps.csr_matrix(mine_d[pp.DISCRETIZATION_MATRICES]["transport"][key])).sum() == 0, key

    def blocks():
        nh, nl = sd_h.num_cells, sd_l.num_cells
        return np.array([[sps.coo_matrix((a, b)) for b in (nh, nl, intf_r.num_cells)] for a in (nh, nl, intf_r.num_cells)],
                        dtype=object)
    A_ref, _ = ref.assemble_matrix_rhs(sd_h, sd_l, intf_r, {}, {}, ref_d, blocks())
    A_mine, _ = mine.assemble_matrix_rhs(sd_h, sd_l, intf_r, {}, {}, mine_d, blocks())
    assert abs(sps.bmat(A_ref) - sps.bmat(A_mine)).sum() == 0
