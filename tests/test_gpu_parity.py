"""Parity of the CUDA path (through the C-ABI, via the pb.Mpfa / pb.Mpsa / pb.Biot operator
API) against (i) the reference's golden outputs, (ii) the oracle on seeded inputs, (iii)
size-independent properties at BASELINE config sizes.  Tolerance: 1e-10 relative to the matrix
max-abs (north_star: "transmissibilities within 1e-10 relative"), solutions 1e-9 in the 2-norm."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from cases import flatten, load_case, max_rel_err
from golden_io import case_names, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _params_mpfa(c):
    return {"second_order_tensor": pb.SecondOrderTensor.from_values(c.raw["K"]), "bc": c.bc,
            "mpfa_eta": c.eta}


@pytest.mark.parametrize("name", case_names("mpfa_"))
def test_mpfa_golden(name):
    c = load_case(name)
    data = pb.initialize_data({}, "flow", _params_mpfa(c))
    d = pb.Mpfa("flow")
    d.discretize(c.g, data)
    out = data[pb.DISCRETIZATION_MATRICES]["flow"]
    assert set(out) == set(c.mats)
    err, key = max_rel_err(c.mats, out)
    assert err < TOL, (key, err)
    data[pb.PARAMETERS]["flow"]["bc_values"] = c.raw["bc_values"]
    A, b = d.assemble_matrix_rhs(c.g, data)
    p = spla.spsolve(sps.csc_matrix(A), b)
    assert np.linalg.norm(p - c.raw["solution"]) <= 1e-9 * np.linalg.norm(c.raw["solution"])


@pytest.mark.parametrize("name", case_names("mpsa_") + case_names("biot_") + case_names("rotbasis_"))
def test_mpsa_biot_golden(name):
    """incl. vectorial boundary conditions in rotated bases (``bc.basis``, a different rotation per face)."""
    c = load_case(name)
    params = {"fourth_order_tensor": pb.FourthOrderTensor.from_values(c.raw["C"]), "bc": c.bc,
              "mpsa_eta": c.eta}
    if c.alpha:
        params["scalar_vector_mappings"] = {
            k: (float(v) if np.ndim(v) == 0 else pb.SecondOrderTensor.from_values(v))
            for k, v in c.alpha.items()}
    data = pb.initialize_data({}, "mech", params)
    d = pb.Biot("mech") if c.alpha else pb.Mpsa("mech")
    d.discretize(c.g, data)
    out = data[pb.DISCRETIZATION_MATRICES]["mech"]
    assert set(flatten(out)) == set(c.mats)
    err, key = max_rel_err(c.mats, out)
    assert err < TOL, (key, err)


def _aniso(nc, rng):
    return pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                                0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))


def _mixed_scalar_bc(g):
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    lab = np.where((x < 1e-10) | (x > 1 - 1e-10), "dir", "neu")
    return pb.BoundaryCondition(g, bf, list(lab))


def _mixed_vector_bc(g):
    bf = g.get_all_boundary_faces()
    bc = pb.BoundaryConditionVectorial(g)
    bot = bf[g.face_centers[2, bf] < 1e-10]
    bc.is_dir[:, bot] = True
    bc.is_neu[:, bot] = False
    side = bf[g.face_centers[0, bf] < 1e-10]
    bc.is_dir[0, side] = True  # roller
    bc.is_neu[0, side] = False
    return bc


@pytest.mark.parametrize("make", [lambda: pb.cart_grid_3d([7, 6, 5], perturb=0.3, seed=3),
                                  lambda: pb.structured_tet_grid([3, 3, 3])])
def test_mpfa_vs_oracle_seeded(make):
    from oracle import fv_oracle as fo
    g = make()
    rng = np.random.default_rng(7)
    k = _aniso(g.num_cells, rng)
    bc = _mixed_scalar_bc(g)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    pb.Mpfa("flow").discretize(g, data)
    ref = fo.mpfa(g, k.values, bc, pb.determine_eta(g))
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < TOL, (key, err)


@pytest.mark.parametrize("make", [lambda: pb.cart_grid_3d([5, 4, 4], perturb=0.3, seed=5),
                                  lambda: pb.structured_tet_grid([2, 3, 2])])
def test_biot_vs_oracle_seeded(make):
    from oracle import fv_oracle as fo
    g = make()
    rng = np.random.default_rng(11)
    nc = g.num_cells
    C = pb.FourthOrderTensor(np.exp(0.5 * rng.standard_normal(nc)), np.exp(0.5 * rng.standard_normal(nc)))
    bc = _mixed_vector_bc(g)
    al = _aniso(nc, rng)
    data = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": bc,
                                           "scalar_vector_mappings": {"flow": al, "t": 0.7}})
    pb.Biot("mech").discretize(g, data)
    ref = flatten(fo.mpsa(g, C.values, bc, pb.determine_eta(g), alpha={"flow": al.values, "t": 0.7}))
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["mech"])
    assert err < TOL, (key, err)


# ---- size-independent properties at config[0] size (32^3) and on ~10^5-cell tets ------------


@pytest.mark.parametrize("make", [lambda: pb.cart_grid_3d([32, 32, 32], perturb=0.2, seed=1),
                                  lambda: pb.structured_tet_grid([26, 26, 26])])
def test_mpfa_linear_pressure_exact(make):
    """MPFA reproduces linear pressure fields exactly for cell-wise constant K
    (tests/numerics/fv/test_mpfa.py:74-137): with p = a.x and Dirichlet data p_b = a.x_f the
    face fluxes equal -n_f.K a, and the reconstructed boundary pressure equals a.x_f."""
    g = make()
    nc = g.num_cells
    Kc = np.array([[2.0, 0.3, 0.1], [0.3, 1.5, 0.2], [0.1, 0.2, 1.2]])
    k = pb.SecondOrderTensor(*[np.full(nc, Kc[i, j]) for i, j in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))])
    bf = g.get_all_boundary_faces()
    bc = pb.BoundaryCondition(g, bf, "dir")
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    d = pb.Mpfa("flow")
    d.discretize(g, data)
    M = data[pb.DISCRETIZATION_MATRICES]["flow"]
    a = np.array([1.0, -2.0, 0.5])
    pc = a @ g.cell_centers
    pb_ = np.zeros(g.num_faces)
    pb_[bf] = a @ g.face_centers[:, bf]
    flux = M["flux"] @ pc + M["bound_flux"] @ pb_
    exact = -(g.face_normals.T @ (Kc @ a))
    assert np.abs(flux - exact).max() <= 1e-10 * np.abs(exact).max()
    trace = M["bound_pressure_cell"] @ pc + M["bound_pressure_face"] @ pb_
    assert np.abs(trace[bf] - pb_[bf]).max() <= 1e-10 * np.abs(pb_).max()
    # gravity-type vector source: p = const, v = K^{-1}-free check: flux = n_f.K v (mpfa.py:1158)
    v = np.array([0.3, -0.2, 1.0])
    vs = M["vector_source"] @ np.tile(v, nc)
    # with p = 0 inside and on the boundary the only driving force is v; constant v on a
    # Dirichlet domain is equivalent to the linear pressure p = v.x up to the boundary data
    pcv = v @ g.cell_centers
    pbv = np.zeros(g.num_faces)
    pbv[bf] = v @ g.face_centers[:, bf]
    zero = M["flux"] @ pcv + M["bound_flux"] @ pbv + vs
    assert np.abs(zero).max() <= 1e-9 * np.abs(vs).max()


def test_mpsa_rigid_and_uniform_strain_cart32():
    """Uniform strain / rigid translation exactness (tests/numerics/fv/test_mpsa.py:213-347) at
    config size 32^3: u = B x on a Dirichlet domain gives tractions n_f.(C : sym B) on every face."""
    g = pb.cart_grid_3d([32, 32, 32], perturb=0.2, seed=2)
    nc = g.num_cells
    mu, lam = 1.3, 0.7
    C = pb.FourthOrderTensor(np.full(nc, mu), np.full(nc, lam))
    bf = g.get_all_boundary_faces()
    bc = pb.BoundaryConditionVectorial(g, bf, "dir")
    data = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": bc})
    pb.Mpsa("mech").discretize(g, data)
    M = data[pb.DISCRETIZATION_MATRICES]["mech"]
    B = np.array([[0.2, 0.1, -0.3], [0.05, -0.1, 0.2], [0.3, 0.0, 0.15]])
    uc = (B @ g.cell_centers).ravel("F")
    ub = np.zeros((3, g.num_faces))
    ub[:, bf] = B @ g.face_centers[:, bf]
    t = M["stress"] @ uc + M["bound_stress"] @ ub.ravel("F")
    sig = mu * (B + B.T) + lam * np.trace(B) * np.eye(3)
    exact = (sig @ g.face_normals).ravel("F")
    assert np.abs(t - exact).max() <= 1e-9 * np.abs(exact).max()
    # rigid translation: zero traction
    uc = np.tile([1.0, 2.0, 3.0], nc)
    ub = np.zeros((3, g.num_faces))
    ub[:, bf] = np.array([[1.0], [2.0], [3.0]])
    t = M["stress"] @ uc + M["bound_stress"] @ ub.ravel("F")
    assert np.abs(t).max() <= 1e-9 * np.abs(exact).max()
    tr = M["bound_displacement_cell"] @ uc + M["bound_displacement_face"] @ ub.ravel("F")
    assert np.abs(tr.reshape(-1, 3)[bf] - [1.0, 2.0, 3.0]).max() <= 1e-10


def test_singular_system_raises_value_error():
    """matrix_operations.py:1487-1490: ValueError('Error in inversion of local linear systems')."""
    g = pb.cart_grid_3d([3, 3, 3])
    k = pb.SecondOrderTensor(np.zeros(g.num_cells))
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": pb.BoundaryCondition(g)})
    with pytest.raises(ValueError, match="inversion of local linear systems"):
        pb.Mpfa("flow").discretize(g, data)


def test_pyramid_raises_assertion_error():
    nodes = np.array([[0, 1, 1, 0, .5], [0, 0, 1, 1, .5], [0, 0, 0, 0, 1.]])
    faces = [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]]
    ind = np.concatenate(faces)
    ptr = np.r_[0, np.cumsum([len(f) for f in faces])]
    fn = sps.csc_matrix((np.ones(ind.size, bool), ind, ptr), shape=(5, 5))
    g = pb.Grid(3, nodes, fn, sps.csc_matrix(np.ones((5, 1))))
    g.set_geometry(np.zeros((3, 5)), np.zeros((3, 5)), np.ones(5), np.zeros((3, 1)), np.ones(1))
    data = pb.initialize_data({}, "flow", {"second_order_tensor": pb.SecondOrderTensor(np.ones(1)),
                                           "bc": pb.BoundaryCondition(g)})
    with pytest.raises(AssertionError):
        pb.Mpfa("flow").discretize(g, data)


def test_spmv_matches_scipy():
    g = pb.cart_grid_3d([24, 20, 16], perturb=0.2)
    rng = np.random.default_rng(0)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": _aniso(g.num_cells, rng),
                                           "bc": _mixed_scalar_bc(g)})
    d = pb.Mpfa("flow")
    d.discretize(g, data)
    data[pb.PARAMETERS]["flow"]["bc_values"] = np.zeros(g.num_faces)
    A, _ = d.assemble_matrix_rhs(g, data)
    x = rng.standard_normal(A.shape[1])
    y = pb.DeviceCsr(A) @ x
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    # ragged / empty rows
    R = sps.random(1000, 700, density=0.01, random_state=1, format="csr")
    R = sps.vstack([R, sps.csr_matrix((5, 700))]).tocsr()
    x = rng.standard_normal(700)
    assert np.abs(pb.DeviceCsr(R) @ x - R @ x).max() <= 1e-12


def test_device_pattern_expansion_matches_host():
    from porepy_b200.fv import DevicePlan, block_expand
    g = pb.structured_tet_grid([4, 3, 3])
    plan = DevicePlan.for_grid(g)
    for which in range(4):
        ip, ix = plan.base_pattern(which)
        for br, bc in ((3, 3), (1, 3), (3, 1)):
            nip, nix = plan.pattern(which, br, bc)
            hip, hix = block_expand(ip, ix, br, bc)
            assert np.array_equal(nip, hip) and np.array_equal(nix, hix)


def test_sharded_equals_unsplit():
    """Two shards discretized one after the other on the GPU reproduce the unsplit matrices
    (common_xpfa_tests.py:832-957)."""
    from porepy_b200 import shard as sh
    g = pb.cart_grid_3d([8, 5, 4], perturb=0.3, seed=9)
    rng = np.random.default_rng(5)
    k = _aniso(g.num_cells, rng)
    bc = _mixed_scalar_bc(g)
    C = pb.FourthOrderTensor(np.exp(0.3 * rng.standard_normal(g.num_cells)), np.ones(g.num_cells))
    vbc = _mixed_vector_bc(g)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    pb.Mpfa("flow").discretize(g, data)
    dm = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc})
    pb.Mpsa("mech").discretize(g, dm)
    ref = dict(data[pb.DISCRETIZATION_MATRICES]["flow"])
    ref.update(dm[pb.DISCRETIZATION_MATRICES]["mech"])
    part = sh.partition_cells(g, 2)
    acc = {}
    shapes = {"flux": ("face", "cell", 1, 1), "bound_flux": ("face", "face", 1, 1),
              "bound_pressure_cell": ("face", "cell", 1, 1), "bound_pressure_face": ("face", "face", 1, 1),
              "vector_source": ("face", "cell", 1, 3), "bound_pressure_vector_source": ("face", "cell", 1, 3),
              "stress": ("face", "cell", 3, 3), "bound_stress": ("face", "face", 3, 3),
              "bound_displacement_cell": ("face", "cell", 3, 3),
              "bound_displacement_face": ("face", "face", 3, 3)}
    for r in range(2):
        s = sh.extract_shard(g, part, r)
        pb.DevicePlan.for_grid(s.grid).set_active_nodes(s.own_node)   # the shard's own interaction regions only
        d1 = pb.initialize_data({}, "flow", {
            "second_order_tensor": pb.SecondOrderTensor.from_values(s.restrict_cell_array(k.values)),
            "bc": sh.restrict_scalar_bc(bc, s)})
        pb.Mpfa("flow").discretize(s.grid, d1)
        d2 = pb.initialize_data({}, "mech", {
            "fourth_order_tensor": pb.FourthOrderTensor.from_values(s.restrict_cell_array(C.values)),
            "bc": sh.restrict_vector_bc(vbc, s)})
        pb.Mpsa("mech").discretize(s.grid, d2)
        loc = dict(d1[pb.DISCRETIZATION_MATRICES]["flow"])
        loc.update(d2[pb.DISCRETIZATION_MATRICES]["mech"])
        for key, m in loc.items():
            gm = s.to_global(m, *shapes[key])
            acc[key] = gm if key not in acc else acc[key] + gm
    for key in ref:
        assert rel_err(ref[key], acc[key]) < 1e-12, key


@pytest.mark.parametrize("make", [lambda: pb.cart_grid_3d([5, 4, 4], perturb=0.3, seed=8),
                                  lambda: pb.structured_tet_grid([2, 2, 3])])
def test_high_contrast_vs_oracle(make):
    """Six orders of magnitude of contrast in permeability and stiffness between neighbouring cells
    (test_mpfa.py:140-251 pattern): the threshold block pivoting must hold the 1e-10 entrywise bar."""
    from oracle import fv_oracle as fo
    g = make()
    rng = np.random.default_rng(21)
    nc = g.num_cells
    kk = np.where(rng.random(nc) < 0.5, 1e-3, 1e3)
    k = pb.SecondOrderTensor(kk)
    bc = _mixed_scalar_bc(g)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
    pb.Mpfa("flow").discretize(g, data)
    ref = fo.mpfa(g, k.values, bc, pb.determine_eta(g))
    err, key = max_rel_err(ref, data[pb.DISCRETIZATION_MATRICES]["flow"])
    assert err < TOL, (key, err)
    mu = np.where(rng.random(nc) < 0.5, 1e-3, 1e3)
    C = pb.FourthOrderTensor(mu, 2.0 * mu)
    vbc = _mixed_vector_bc(g)
    dm = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc})
    pb.Mpsa("mech").discretize(g, dm)
    refm = fo.mpsa(g, C.values, vbc, pb.determine_eta(g))
    err, key = max_rel_err(refm, dm[pb.DISCRETIZATION_MATRICES]["mech"])
    assert err < TOL, (key, err)


def test_device_side_flow_system_and_solve():
    """``discretize`` leaves device-resident lazy matrices; ``assemble_matrix_rhs`` then forms A = div @ flux and b on
    the device FROM THOSE MATRICES (no download) and equals the host products of fv_elliptic.py:67-112; two
    keywords on one grid do not see each other's values; BiCGStab on the device matrix reproduces the direct solve."""
    from porepy_b200 import krylov as kr
    from porepy_b200.sparse import LazyCsr
    g = pb.structured_tet_grid([6, 5, 4])
    rng = np.random.default_rng(3)
    k = _aniso(g.num_cells, rng)
    bc = _mixed_scalar_bc(g)
    bf = g.get_all_boundary_faces()
    bv = np.zeros(g.num_faces)
    bv[bf] = rng.random(bf.size)
    vs = rng.standard_normal(3 * g.num_cells)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc, "bc_values": bv,
                                           "vector_source": vs})
    d = pb.Mpfa("flow")
    d.discretize(g, data)
    mats = data[pb.DISCRETIZATION_MATRICES]["flow"]
    assert all(isinstance(m, LazyCsr) and not m.on_host for m in mats.values())
    # a second keyword on the same grid (shared plan) with another permeability, discretized AFTERWARDS
    other = pb.initialize_data({}, "fourier", {"second_order_tensor": pb.SecondOrderTensor(7.0 * np.ones(g.num_cells)),
                                               "bc": bc, "bc_values": bv})
    pb.Mpfa("fourier").discretize(g, other)
    before = dict(LazyCsr.downloads)
    A_lazy, b_dev = d.assemble_matrix_rhs(g, data)
    A_dev = A_lazy.device_csr
    assert A_dev is not None and LazyCsr.downloads == before and not mats["flux"].on_host
    x = rng.standard_normal(g.num_cells)
    y_dev = A_dev @ x
    # host products from downloaded copies of the SAME stored matrices
    div = g.divergence(dim=1)
    A_host = (div @ sps.csr_matrix(mats["flux"])).tocsr()
    b_host = -div @ (mats["bound_flux"] @ bv) - div @ (mats["vector_source"] @ vs)
    assert mats["flux"].on_host
    assert rel_err(A_host, A_dev.to_scipy()) < 1e-13
    assert rel_err(A_host, A_lazy) < 1e-13            # the lazy system downloads on touch
    assert np.abs(b_dev - b_host).max() <= 1e-12 * np.abs(b_host).max()
    assert np.abs(y_dev - A_host @ x).max() <= 1e-12 * np.abs(A_host @ x).max()
    # once touched by the host the matrices are used from the host (they may have been edited there)
    A2, b2 = d.assemble_matrix_rhs(g, data)
    assert getattr(A2, "device_csr", None) is None and rel_err(A_host, A2) < 1e-13
    # solve on the device matrix
    loc = kr.LocalSystem(0, 1, np.arange(g.num_cells), np.zeros(0, np.int64), A_dev, [0], [np.zeros(0, np.int64)])
    xs, info = kr.solve_local(loc, b_dev, diag_own=A_dev.diagonal(), tol=1e-11)
    ref = spla.spsolve(sps.csc_matrix(A_host), b_host)
    assert info["converged"] and np.linalg.norm(xs.cpu().numpy() - ref) <= 1e-8 * np.linalg.norm(ref)


def test_device_side_mechanics_system():
    """A = div_nd @ stress and b = -div_nd @ (bound_stress @ bc) + source assembled on the device equal
    the host products of mpsa.py:486-529."""
    g = pb.cart_grid_3d([6, 5, 4], perturb=0.3, seed=2)
    rng = np.random.default_rng(5)
    nc = g.num_cells
    C = pb.FourthOrderTensor(np.exp(0.4 * rng.standard_normal(nc)), np.exp(0.4 * rng.standard_normal(nc)))
    vbc = _mixed_vector_bc(g)
    bv = np.zeros((3, g.num_faces))
    bf = g.get_all_boundary_faces()
    bv[:, bf] = rng.random((3, bf.size))
    src = rng.standard_normal(3 * nc)
    data = pb.initialize_data({}, "mech", {"fourth_order_tensor": C, "bc": vbc,
                                           "bc_values": bv.ravel("F"), "source": src})
    d = pb.Mpsa("mech")
    d.discretize(g, data)
    A_dev, b_dev = d.assemble_matrix_rhs_device(g, data)
    mats = data[pb.DISCRETIZATION_MATRICES]["mech"]
    div = g.divergence(dim=3)
    A_host = (div @ sps.csr_matrix(mats["stress"])).tocsr()
    b_host = -div @ (mats["bound_stress"] @ bv.ravel("F")) + src
    assert A_dev.shape == A_host.shape
    assert rel_err(A_host, A_dev.to_scipy()) < 1e-13
    assert np.abs(b_dev - b_host).max() <= 1e-12 * np.abs(b_host).max()
    x = rng.standard_normal(3 * nc)
    assert np.abs(A_dev @ x - A_host @ x).max() <= 1e-12 * np.abs(A_host @ x).max()
    s, q = A_dev.checksum()
    assert abs(s - A_host.data.sum()) <= 1e-9 * np.abs(A_host.data).sum() and abs(q - (A_host.data ** 2).sum()) <= 1e-10 * q


def test_sharded_system_rows_on_the_device():
    """Two shards of one mesh, each on its own plan restricted to its own nodes (``set_active_nodes``): the rows of
    the own cells of the device-assembled flow system are the rows of the unsplit system (columns [own | ghost])."""
    from porepy_b200 import shard as sh
    g = pb.structured_tet_grid([5, 4, 4])
    rng = np.random.default_rng(9)
    k = _aniso(g.num_cells, rng)
    bc = _mixed_scalar_bc(g)
    bf = g.get_all_boundary_faces()
    bv = np.zeros(g.num_faces)
    bv[bf] = rng.random(bf.size)
    data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc, "bc_values": bv})
    d = pb.Mpfa("flow")
    d.discretize(g, data)
    A, b = d.assemble_matrix_rhs(g, data)
    A = sps.csr_matrix(A)
    part = sh.partition_cells(g, 2)
    for r in range(2):
        s = sh.extract_shard(g, part, r)
        n_own = int(s.own_cell.sum())
        plan = pb.DevicePlan.for_grid(s.grid)
        plan.set_active_nodes(s.own_node)
        # rank 0 restricts the permeability on the host, rank 1 hands over the GLOBAL tensor (device-side gather)
        if r == 1:
            plan.set_cell_map(s.cells, g.num_cells)
        kl = k if r == 1 else pb.SecondOrderTensor.from_values(s.restrict_cell_array(k.values))
        dl = pb.initialize_data({}, "flow", {
            "second_order_tensor": kl,
            "bc": sh.restrict_scalar_bc(bc, s), "bc_values": bv[s.faces], "mpfa_eta": pb.determine_eta(g)})
        dd = pb.Mpfa("flow")
        dd.discretize(s.grid, dl)
        a_dev, b_loc = dd.assemble_matrix_rhs_device(s.grid, dl)
        rows = a_dev.truncate_rows(n_own).to_scipy()
        ref = A[s.cells[:n_own]][:, s.cells]
        assert abs(ref - rows).max() <= 1e-12 * abs(A).max()
        assert np.abs(b_loc[:n_own] - b[s.cells[:n_own]]).max() <= 1e-12 * np.abs(b).max()
