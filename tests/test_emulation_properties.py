"""Domain properties of the node routines (host build, tests/emu) on grids and boundary
conditions beyond the golden fixtures, following the reference's own exactness tests:
linear pressure / uniform strain reproduction with Dirichlet, Neumann and Robin data
(tests/numerics/fv/test_mpfa.py:74-137,866-1040; test_mpsa.py:213-347,1189-1323), hydrostatic
equilibrium with a vector source (test_mpfa.py gravity tests), consistency of the Biot coupling
terms (test_biot.py).  No GPU; the same source (mpfa_node / mpsa_node) is what the kernels run."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import porepy_b200 as pb
from emu_binding import EmuPlan

KC = np.array([[2.0, 0.3, 0.1], [0.3, 1.5, 0.2], [0.1, 0.2, 1.2]])


def jittered_tets(n=4, amp=0.15, seed=0):
    """Structured tetrahedra with randomly displaced interior nodes (general simplices)."""
    g = pb.structured_tet_grid([n, n, n])
    rng = np.random.default_rng(seed)
    nodes = g.nodes.copy()
    interior = np.all((nodes > 1e-9) & (nodes < 1 - 1e-9), axis=0)
    nodes[:, interior] += amp / n * (rng.random((3, interior.sum())) - 0.5)
    # cell -> nodes from the face-node / cell-face incidence
    fn = sps.csc_matrix(g.face_nodes)
    cf = sps.csc_matrix(g.cell_faces)
    cn = np.zeros((g.num_cells, 4), dtype=np.int64)
    for c in range(g.num_cells):
        faces = cf.indices[cf.indptr[c]:cf.indptr[c + 1]]
        cn[c] = np.unique(np.concatenate([fn.indices[fn.indptr[f]:fn.indptr[f + 1]] for f in faces]))
    return pb.tet_grid_from_cells(nodes, cn)


GRIDS = [lambda: pb.cart_grid_3d([5, 4, 3], perturb=0.3, seed=1), lambda: jittered_tets(3, seed=2)]


def perm_values(nc):
    v = np.zeros((3, 3, nc))
    v[:] = KC[:, :, None]
    return v


def boundary_sets(g):
    bf = g.get_all_boundary_faces()
    x = g.face_centers[0, bf]
    z = g.face_centers[2, bf]
    west, top = bf[x < 1e-9], bf[z > 1 - 1e-9]
    rest = np.setdiff1d(bf, np.r_[west, top])
    return bf, west, top, rest


def outward_sign(g):
    """+1 / -1 per boundary face: face normal points out of / into the domain."""
    cf = sps.csr_matrix(g.cell_faces)
    return np.asarray(cf.sum(axis=1)).ravel()


@pytest.mark.parametrize("make", GRIDS)
def test_mpfa_linear_pressure_with_dirichlet_neumann_and_robin_data(make):
    """p = a.x is reproduced exactly when every boundary type is fed with its exact datum:
    Dirichlet p_b, Neumann outward flux, Robin flux + w A p (mpfa.py:869-887,1516-1547)."""
    g = make()
    nc, nf = g.num_cells, g.num_faces
    bf, dirf, robf, neuf = boundary_sets(g)
    codes = np.zeros(nf, np.uint8)
    codes[dirf], codes[neuf], codes[robf] = 1, 2, 3
    w = 1.7
    robw = np.full(nf, w)
    out = EmuPlan(g).mpfa(perm_values(nc), codes, robw, pb.determine_eta(g))
    a = np.array([1.0, -2.0, 0.5])
    pc = a @ g.cell_centers
    exact_flux = -(g.face_normals.T @ (KC @ a))      # along the face normal
    sgn = outward_sign(g)
    bcv = np.zeros(nf)
    bcv[dirf] = a @ g.face_centers[:, dirf]
    bcv[neuf] = exact_flux[neuf] * sgn[neuf]          # outward flux
    bcv[robf] = exact_flux[robf] * sgn[robf] + w * g.face_areas[robf] * (a @ g.face_centers[:, robf])
    div = g.divergence(1)
    A = div @ out["flux"]
    b = -div @ (out["bound_flux"] @ bcv)
    p = spla.spsolve(sps.csc_matrix(A), b)
    assert np.abs(p - pc).max() <= 1e-9 * np.abs(pc).max()
    flux = out["flux"] @ p + out["bound_flux"] @ bcv
    assert np.abs(flux - exact_flux).max() <= 1e-9 * np.abs(exact_flux).max()
    trace = out["bound_pressure_cell"] @ p + out["bound_pressure_face"] @ bcv
    pf = a @ g.face_centers
    assert np.abs(trace[bf] - pf[bf]).max() <= 1e-9 * np.abs(pf).max()


@pytest.mark.parametrize("make", GRIDS)
def test_mpfa_hydrostatic_equilibrium(make):
    """grad p = v (vector source) gives zero flux for any K, also across Neumann boundaries
    (mpfa.py:1158-1307); the reconstructed boundary pressure carries the vector-source part."""
    g = make()
    nc, nf = g.num_cells, g.num_faces
    rng = np.random.default_rng(5)
    K = perm_values(nc) * (1 + rng.random(nc))          # cell-wise different, same principal axes
    bf, dirf, top, rest = boundary_sets(g)
    codes = np.zeros(nf, np.uint8)
    codes[dirf], codes[np.r_[top, rest]] = 1, 2
    out = EmuPlan(g).mpfa(K, codes, None, pb.determine_eta(g))
    v = np.array([0.0, 0.3, -1.0])
    pc = v @ g.cell_centers
    bcv = np.zeros(nf)
    bcv[dirf] = v @ g.face_centers[:, dirf]             # Neumann data: zero flux
    vs = np.tile(v, nc)
    flux = out["flux"] @ pc + out["bound_flux"] @ bcv + out["vector_source"] @ vs
    scale = np.abs(out["vector_source"] @ vs).max()
    assert np.abs(flux).max() <= 1e-9 * scale
    trace = (out["bound_pressure_cell"] @ pc + out["bound_pressure_face"] @ bcv
             + out["bound_pressure_vector_source"] @ vs)
    pf = v @ g.face_centers
    assert np.abs(trace[bf] - pf[bf]).max() <= 1e-9 * max(1.0, np.abs(pf).max())


def test_mpfa_reduces_to_two_point_flux_on_orthogonal_grid():
    """Isotropic K on an orthogonal grid: MPFA-O degenerates to TPFA with harmonic averaging
    (tutorial flux_discretizations) and div @ flux is symmetric."""
    g = pb.cart_grid_3d([4, 3, 3])
    nc, nf = g.num_cells, g.num_faces
    rng = np.random.default_rng(3)
    kiso = np.exp(rng.standard_normal(nc))
    K = np.zeros((3, 3, nc))
    K[0, 0] = K[1, 1] = K[2, 2] = kiso
    bf = g.get_all_boundary_faces()
    codes = np.zeros(nf, np.uint8)
    codes[bf] = 1
    out = EmuPlan(g).mpfa(K, codes, None, pb.determine_eta(g))
    cf = sps.csc_matrix(g.cell_faces)
    fc = sps.csr_matrix(g.cell_faces)
    tp = sps.lil_matrix((nf, nc))
    for f in range(nf):
        cells = fc.indices[fc.indptr[f]:fc.indptr[f + 1]]
        sg = fc.data[fc.indptr[f]:fc.indptr[f + 1]]
        half = []
        for c in cells:
            d = np.abs((g.face_centers[:, f] - g.cell_centers[:, c]) @ g.face_normals[:, f]) / g.face_areas[f]
            half.append(kiso[c] * g.face_areas[f] / d)
        if len(cells) == 2:
            t = half[0] * half[1] / (half[0] + half[1])
            tp[f, cells[0]], tp[f, cells[1]] = sg[0] * t, sg[1] * t
        else:
            tp[f, cells[0]] = sg[0] * half[0]
    diff = (out["flux"] - sps.csr_matrix(tp)).toarray()
    assert np.abs(diff).max() <= 1e-12 * np.abs(tp.toarray()).max()
    A = (g.divergence(1) @ out["flux"]).toarray()
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
    assert cf.shape == (nf, nc)


def lame_stiffness(nc, mu, lam):
    return pb.FourthOrderTensor(np.full(nc, mu), np.full(nc, lam)).values


@pytest.mark.parametrize("make", GRIDS)
def test_mpsa_uniform_strain_with_mixed_boundary_data(make):
    """u = B x is reproduced exactly with Dirichlet data and exact tractions on one boundary
    plane (component-wise mixed: rollers); rigid motions give zero traction.  Traction data on
    two planes meeting in an edge is NOT exact in the reference either: there the asymmetric
    part of the stress is dropped by _eliminate_ncasym (mpsa.py:1932-2000)."""
    g = make()
    nc, nf = g.num_cells, g.num_faces
    mu, lam = 1.3, 0.7
    bf, west, top, rest = boundary_sets(g)
    codes = np.zeros((3, nf), np.uint8)
    codes[:, np.r_[west, rest]] = 1
    codes[0, top], codes[1, top], codes[2, top] = 1, 2, 2   # rollers: u_x fixed, free in y, z
    out = EmuPlan(g).mpsa(lame_stiffness(nc, mu, lam), codes, None, pb.determine_eta(g))
    B = np.array([[0.2, 0.1, -0.3], [0.05, -0.1, 0.2], [0.3, 0.0, 0.15]])
    sig = mu * (B + B.T) + lam * np.trace(B) * np.eye(3)
    exact_t = sig @ g.face_normals                        # along the face normal, (3, nf)
    sgn = outward_sign(g)
    uf = B @ g.face_centers
    bcv = np.zeros((3, nf))
    for i in range(3):
        d = codes[i] == 1
        n = codes[i] == 2
        bcv[i, d] = uf[i, d]
        bcv[i, n] = exact_t[i, n] * sgn[n]
    div = g.divergence(3)
    A = div @ out["stress"]
    b = -div @ (out["bound_stress"] @ bcv.ravel("F"))
    u = spla.spsolve(sps.csc_matrix(A), b)
    uc = (B @ g.cell_centers).ravel("F")
    assert np.abs(u - uc).max() <= 1e-8 * np.abs(uc).max()
    t = out["stress"] @ u + out["bound_stress"] @ bcv.ravel("F")
    assert np.abs(t - exact_t.ravel("F")).max() <= 1e-8 * np.abs(exact_t).max()
    tr = out["bound_displacement_cell"] @ u + out["bound_displacement_face"] @ bcv.ravel("F")
    assert np.abs(tr.reshape(-1, 3)[bf] - uf.T[bf]).max() <= 1e-8 * np.abs(uf).max()
    # rigid translation + infinitesimal rotation: zero traction everywhere
    R = np.array([[0.0, -0.2, 0.1], [0.2, 0.0, -0.3], [-0.1, 0.3, 0.0]])
    ur = (R @ g.cell_centers + np.array([[1.0], [2.0], [3.0]])).ravel("F")
    ub = np.zeros((3, nf))
    for i in range(3):
        d = codes[i] == 1
        ub[i, d] = (R @ g.face_centers + np.array([[1.0], [2.0], [3.0]]))[i, d]
    t = out["stress"] @ ur + out["bound_stress"] @ ub.ravel("F")
    assert np.abs(t).max() <= 1e-8 * np.abs(exact_t).max()


@pytest.mark.parametrize("make", GRIDS)
def test_biot_coupling_terms_are_consistent(make):
    """div u of a linear displacement field is tr(B) |K| (biot.py:1054-1135), and a uniform
    pressure exerts a force alpha p n_f on every face: scalar_gradient @ p = alpha p n_f
    relative to the zero-displacement state on a clamped domain (biot.py:880-1038)."""
    g = make()
    nc, nf = g.num_cells, g.num_faces
    mu, lam, alpha = 1.1, 0.9, 0.8
    bf = g.get_all_boundary_faces()
    codes = np.zeros((3, nf), np.uint8)
    codes[:, bf] = 1
    out = EmuPlan(g).mpsa(lame_stiffness(nc, mu, lam), codes, None, pb.determine_eta(g), alpha={"flow": alpha})
    B = np.array([[0.2, 0.1, -0.3], [0.05, -0.1, 0.2], [0.3, 0.0, 0.15]])
    uc = (B @ g.cell_centers).ravel("F")
    ub = np.zeros((3, nf))
    ub[:, bf] = B @ g.face_centers[:, bf]
    dv = out["displacement_divergence"]["flow"] @ uc + out["boundary_displacement_divergence"]["flow"] @ ub.ravel("F")
    assert np.abs(dv - alpha * np.trace(B) * g.cell_volumes).max() <= 1e-9 * g.cell_volumes.max()
    # uniform pressure on a clamped body: total face force = stress @ u + scalar_gradient @ p with
    # u = 0 solves the momentum balance (closed-surface integral of n vanishes on every cell)
    p = np.full(nc, 2.5)
    force = out["scalar_gradient"]["flow"] @ p
    div = g.divergence(3)
    assert np.abs(div @ force).max() <= 1e-9 * np.abs(force).max()
    # the pressure contribution to the boundary displacement vanishes for clamped boundaries
    tr = out["bound_displacement_pressure"]["flow"] @ p
    assert np.abs(tr.reshape(-1, 3)[bf]).max() <= 1e-9


@pytest.mark.parametrize("make", GRIDS)
def test_node_routines_match_the_oracle_on_seeded_mixed_problems(make):
    """Heterogeneous tensors, Neumann planes meeting in edges and corners (the _eliminate_ncasym
    branch), Robin faces with a full weight matrix, two coupling tensors: matrices against the
    NumPy restatement of the reference (oracle/, pinned to the reference's outputs)."""
    from cases import flatten, max_rel_err, scalar_codes, vector_codes
    from oracle import fv_oracle as fo
    g = make()
    nc, nf = g.num_cells, g.num_faces
    rng = np.random.default_rng(17)
    bf, west, top, rest = boundary_sets(g)
    lab = np.array(["neu"] * bf.size, dtype=object)
    lab[np.isin(bf, west)] = "dir"
    lab[np.isin(bf, top)] = "rob"
    # flow
    k = pb.SecondOrderTensor(1 + rng.random(nc), 1 + rng.random(nc), 1 + rng.random(nc),
                             0.3 * rng.random(nc), 0.3 * rng.random(nc), 0.3 * rng.random(nc))
    bc = pb.BoundaryCondition(g, bf, list(lab))
    bc.robin_weight = 0.5 + rng.random(nf)
    eta = pb.determine_eta(g)
    p = EmuPlan(g)
    out = p.mpfa(k.values, scalar_codes(bc, nf), bc.robin_weight, eta)
    err, key = max_rel_err(fo.mpfa(g, k.values, bc, eta), out)
    assert err < 1e-10, (key, err)
    # mechanics + coupling
    C = pb.FourthOrderTensor(np.exp(0.5 * rng.standard_normal(nc)), np.exp(0.5 * rng.standard_normal(nc)))
    vbc = pb.BoundaryConditionVectorial(g, bf, list(lab))
    w = rng.random((3, 3, nf))
    vbc.robin_weight = w + np.transpose(w, (1, 0, 2)) + 3 * np.eye(3)[:, :, None]
    al = {"flow": k.values, "t": 0.7}
    out = p.mpsa(C.values, vector_codes(vbc, 3, nf), vbc.robin_weight, eta, alpha=al)
    err, key = max_rel_err(flatten(fo.mpsa(g, C.values, vbc, eta, alpha=al)), out)
    assert err < 1e-10, (key, err)
