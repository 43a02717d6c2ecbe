"""Host build (1-thread team) of the CUDA node routines + the C++ plan builder against the
reference's golden outputs.  Checks the reduced continuity-point formulation, the plan indexing
and the pattern/scatter maps without a GPU.  The emulation library is test infrastructure
(tests/emu); the product never loads it."""
import numpy as np
import pytest
import scipy.sparse as sps

from cases import load_case, max_rel_err, scalar_codes, vector_codes
from emu_binding import EmuPlan
from golden_io import case_names

TOL = 1e-10  # north_star tolerance; observed <= 2e-15 (Delaunay slivers ~1e-12)


@pytest.mark.parametrize("name", case_names("mpfa_"))
def test_mpfa_node_routine(name):
    c = load_case(name)
    p = EmuPlan(c.g)
    out = p.mpfa(c.raw["K"], scalar_codes(c.bc, c.g.num_faces), c.bc.robin_weight, c.eta)
    err, key = max_rel_err(c.mats, out)
    assert err < TOL, (key, err)
    # the solved problem (incl. the kappa = 1e+-6 contrast case): relative 2-norm
    import scipy.sparse.linalg as spla
    div = c.g.divergence(1)
    A = div @ out["flux"]
    b = -div @ (out["bound_flux"] @ c.raw["bc_values"])
    sol = spla.spsolve(sps.csc_matrix(A), b)
    assert np.linalg.norm(sol - c.raw["solution"]) <= 1e-9 * np.linalg.norm(c.raw["solution"])


@pytest.mark.parametrize("name", case_names("mpsa_") + case_names("biot_"))
def test_mpsa_node_routine(name):
    c = load_case(name)
    nd = c.g.dim
    p = EmuPlan(c.g)
    out = p.mpsa(c.raw["C"], vector_codes(c.bc, nd, c.g.num_faces), c.bc.robin_weight[:nd, :nd],
                 c.eta, alpha=c.alpha or None)
    err, key = max_rel_err(c.mats, out)
    assert err < TOL, (key, err)


def test_pyramid_rejected_by_plan_builder():
    from porepy_b200.grid import Grid
    nodes = np.array([[0, 1, 1, 0, .5], [0, 0, 1, 1, .5], [0, 0, 0, 0, 1.]])
    faces = [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]]
    ind = np.concatenate(faces)
    ptr = np.r_[0, np.cumsum([len(f) for f in faces])]
    fn = sps.csc_matrix((np.ones(ind.size, bool), ind, ptr), shape=(5, 5))
    g = Grid(3, nodes, fn, sps.csc_matrix(np.ones((5, 1))))
    g.set_geometry(np.zeros((3, 5)), np.zeros((3, 5)), np.ones(5), np.zeros((3, 1)), np.ones(1))
    with pytest.raises(AssertionError):
        EmuPlan(g)


def test_patterns_are_structural_supersets():
    """Every non-zero of the reference lies inside the structural pattern."""
    c = load_case("mpfa_cart3d")
    p = EmuPlan(c.g)
    ip, ix = p.pat[0]
    pat = sps.csr_matrix((np.ones(ix.size), ix, ip), shape=c.mats["flux"].shape)
    ref = c.mats["flux"].copy()
    ref.data = (abs(ref.data) > 1e-14).astype(float)
    assert (ref - ref.multiply(pat)).nnz == 0


@pytest.mark.parametrize("name", case_names("rotbasis_"))
def test_mpsa_node_routine_with_rotated_boundary_bases(name):
    """Vectorial boundary conditions given in a rotated frame, a different rotation on every face
    (bc.basis; _fvutils.py:765-945), Dirichlet / Neumann / Robin mixed per rotated component, with
    and without Biot coupling: golden outputs of the reference."""
    c = load_case(name)
    nd = c.g.dim
    out = EmuPlan(c.g).mpsa(c.raw["C"], vector_codes(c.bc, nd, c.g.num_faces), c.bc.robin_weight[:nd, :nd],
                            c.eta, alpha=c.alpha or None, basis=c.raw["bc_basis"])
    err, key = max_rel_err(c.mats, out)
    assert err < TOL, (key, err)
