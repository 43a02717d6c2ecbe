"""``pb.DifferentiableTpfa`` (reference numerics/fv/tpfa.py:281-760 + the AD transmissibility expression of
models/constitutive_laws.py:1544-1583) against ``difftpfa_*`` fixtures written from the unmodified reference
(tools/make_golden.py ``case_diff_tpfa``): the helper matrices entry by entry, and the fused evaluation -- face
transmissibilities, half-face values and the Jacobian dT_f/dk_c -- against the reference's own AdArray chain.
CPU: host build of the per-face routine; GPU: ``pb_tpfa_diff`` through the C ABI."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

import porepy_b200 as pb
from porepy_b200 import fv
from golden_io import case_names, load_case

CASES = case_names("difftpfa_")


def _grid(c):
    g = c.g
    g.tags["tip_faces"] = np.asarray(c.raw["tip_faces"], bool)
    g.tags["domain_boundary_faces"] = np.asarray(c.raw["domain_boundary_faces"], bool)
    return g


def _same(ref, got, tol=0.0):
    ref, got = sps.csr_matrix(ref), sps.csr_matrix(got)
    assert ref.shape == got.shape
    d = abs(ref - got)
    return (d.max() if d.nnz else 0.0) <= tol * max(abs(ref).max() if ref.nnz else 1.0, 1e-300)


@pytest.mark.parametrize("name", CASES)
def test_helper_matrices_equal_the_reference(name):
    c = load_case(name)
    g = _grid(c)
    dt = pb.DifferentiableTpfa()
    n, d_vec, dist = dt.half_face_geometry_matrices([g])
    assert _same(c.mats["n"], n) and _same(c.mats["d_vec"], d_vec)
    assert np.array_equal(dist, c.raw["dist"])
    assert _same(c.mats["hf_to_f_signed"], dt.half_face_map([g], to_entity="faces", with_sign=True))
    assert _same(c.mats["c_to_hf"], dt.half_face_map([g], to_entity="half_faces", from_entity="cells"))
    assert _same(c.mats["c3_to_hf3"], dt.half_face_map([g], from_entity="cells", to_entity="half_faces", dimensions=(3, 3)))
    assert _same(c.mats["hf3_to_f"], dt.half_face_map([g], from_entity="half_faces", to_entity="faces",
                                                       dimensions=(1, 3), with_sign=True))
    assert _same(c.mats["face_pairing"], dt.face_pairing_from_cell_array([g]))
    assert _same(c.mats["nd_to_3d_cells_2"], dt.nd_to_3d([g], 2))
    assert _same(c.mats["nd_to_3d_faces_3"], dt.nd_to_3d([g], 3, "faces"))
    assert np.array_equal(dt.boundary_sign([g]), c.raw["boundary_sign"])
    assert np.array_equal(dt.internal_boundary_filter([g]).astype(float), c.raw["internal_boundary_filter"])
    assert np.array_equal(dt.tip_filter([g]).astype(float), c.raw["tip_filter"])
    # two subdomains: block-diagonal concatenation (tpfa.py:371-399)
    two = dt.half_face_map([g, g], to_entity="faces", with_sign=True)
    one = sps.csr_matrix(c.mats["hf_to_f_signed"])
    assert _same(sps.block_diag([one, one]), two)


TOL = 1e-11   # relative to the largest entry; Delaunay slivers have half-face values of opposite sign that cancel in
#               the harmonic sum, where the device's fused multiply-adds round differently from NumPy (observed 1e-13)


def _check_fused(c, g):
    T, jac, t_hf = pb.DifferentiableTpfa().transmissibility(g, c.raw["k_c"])
    assert np.abs(t_hf - c.raw["t_hf"]).max() <= TOL * np.abs(c.raw["t_hf"]).max()
    assert np.abs(T - c.raw["T_f"]).max() <= TOL * np.abs(c.raw["T_f"]).max()
    assert _same(c.mats["dT_dk"], jac, 10 * TOL)
    # chain rule with a permeability Jacobian: k_c = k0 * exp(p_cell), dk/dp is 9 entries per cell
    nc = g.num_cells
    kj = sps.csr_matrix((c.raw["k_c"], (np.arange(9 * nc), np.repeat(np.arange(nc), 9))), shape=(9 * nc, nc))
    _, jac_p, _ = pb.DifferentiableTpfa().transmissibility(g, c.raw["k_c"], k_jac=kj)
    assert _same(sps.csr_matrix(c.mats["dT_dk"]) @ kj, jac_p, 10 * TOL)
    # T is homogeneous of degree one in k: dT/dk . k = T
    assert np.abs(jac @ c.raw["k_c"] - T).max() <= 10 * TOL * np.abs(T).max()


@pytest.mark.parametrize("name", CASES)
def test_fused_evaluation_host_build(name, monkeypatch):
    from emu_binding import EmuBackedFaceGrid
    monkeypatch.setattr(fv, "FaceGrid", EmuBackedFaceGrid)
    c = load_case(name)
    _check_fused(c, _grid(c))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_evaluation_gpu(name):
    c = load_case(name)
    _check_fused(c, _grid(c))
