"""NumPy restatement of the reference's two-point flux approximation and first-order upwinding
(``Tpfa.discretize``, reference src/porepy/numerics/fv/tpfa.py:40-280, and
``Upwind.discretize``, src/porepy/numerics/fv/upwind.py:150-300).

TEST INFRASTRUCTURE ONLY, like oracle/fv_oracle.py: the checker for the kernels of the next scope
row (not built yet).  Pinned against the unmodified reference by tests/golden/next_*.npz
(tools/make_golden.py) in tests/test_oracle_vs_golden.py.  Periodic faces are not restated."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


def _face_cell_pairs(g):
    cf = sps.coo_matrix(g.cell_faces)
    return cf.row.astype(np.int64), cf.col.astype(np.int64), np.asarray(cf.data, dtype=np.float64)


def tpfa(g, k_values: np.ndarray, bc, ambient_dimension: int | None = None) -> dict:
    """Two-point fluxes with harmonic averaging of the half-transmissibilities
    t_{K,f} = n_f^T K_K d / |d|^2, d = x_f - x_K (tpfa.py:162-180); Dirichlet faces keep the half
    transmissibility, Neumann faces carry the prescribed flux (tpfa.py:184-214); pressure trace and
    vector-source terms as tpfa.py:219-279.  Keys as for MPFA."""
    nf, nc = g.num_faces, g.num_cells
    vdim = g.dim if ambient_dimension is None else ambient_dimension
    fi, ci, sgn = _face_cell_pairs(g)
    n = g.face_normals[:, fi] * sgn
    d = g.face_centers[:, fi] - g.cell_centers[:, ci]
    nk = np.einsum("ijc,jc->ic", k_values[:, :, ci], n)          # (K n) per half face (K symmetric)
    t_half = np.einsum("ic,ic->c", nk, d) / np.einsum("ic,ic->c", d, d)
    t = 1.0 / np.bincount(fi, weights=1.0 / t_half, minlength=nf)
    t_full = t.copy()
    internal = np.asarray(getattr(bc, "is_internal", np.zeros(nf, bool)), bool)
    is_dir = np.asarray(bc.is_dir, bool) & ~internal
    is_neu = np.asarray(bc.is_neu, bool) | internal
    bnd = g.get_all_boundary_faces()
    t_b = np.zeros(nf)
    t_b[is_dir] = -t[is_dir]
    t_b[is_neu] = 1.0
    t = t.copy()
    t[is_neu] = 0.0
    flux = sps.coo_matrix((t[fi] * sgn, (fi, ci)), shape=(nf, nc)).tocsr()
    sgn_bnd = np.asarray(sps.csr_matrix(g.cell_faces)[bnd].sum(axis=1)).ravel()
    bound_flux = sps.coo_matrix((t_b[bnd] * sgn_bnd, (bnd, bnd)), shape=(nf, nf)).tocsr()
    neu_raw, dir_raw = np.asarray(bc.is_neu, bool), np.asarray(bc.is_dir, bool)
    v_face = np.zeros(nf)
    v_face[dir_raw] = 1.0
    v_face[neu_raw] = -1.0 / t_full[neu_raw]
    v_cell = np.where(neu_raw[fi], 1.0, 0.0)
    bound_pressure_cell = sps.coo_matrix((v_cell, (fi, ci)), shape=(nf, nc)).tocsr()
    bound_pressure_face = sps.diags(v_face).tocsr()
    rows = np.repeat(fi, vdim)
    cols = (ci[:, None] * vdim + np.arange(vdim)).ravel()
    vals = ((t[fi] * sgn) * d[:vdim]).T.ravel()
    vector_source = sps.coo_matrix((vals, (rows, cols)), shape=(nf, nc * vdim)).tocsr()
    vals = np.where(neu_raw[fi], d[:vdim], 0.0).T.ravel()
    bpvs = sps.coo_matrix((vals, (rows, cols)), shape=(nf, nc * vdim)).tocsr()
    return {"flux": flux, "bound_flux": bound_flux, "bound_pressure_cell": bound_pressure_cell,
            "bound_pressure_face": bound_pressure_face, "vector_source": vector_source,
            "bound_pressure_vector_source": bpvs}


def upwind(g, darcy_flux: np.ndarray, bc) -> dict:
    """First-order upwinding (upwind.py:232-300): per face the upstream cell by the sign of the
    flux (zero counts as positive); Neumann faces and Dirichlet INFLOW faces are removed from the
    matrix and appear in the two boundary matrices.  Keys: upwind, bound_transport_dir,
    bound_transport_neu."""
    nf, nc = g.num_faces, g.num_cells
    fi, ci, sgn = _face_cell_pairs(g)
    cfd = -np.ones((2, nf), dtype=np.int64)     # cell on the + side / - side of every face
    cfd[0, fi[sgn > 0]] = ci[sgn > 0]
    cfd[1, fi[sgn < 0]] = ci[sgn < 0]
    s = np.sign(darcy_flux)
    pos = s >= 0
    up = np.where(pos, cfd[0], cfd[1])
    is_neu, is_dir = np.asarray(bc.is_neu, bool), np.asarray(bc.is_dir, bool)
    inflow = is_dir & ((pos & (cfd[0] < 0)) | (~pos & (cfd[1] < 0)))
    keep = ~(is_neu | inflow)
    rows = np.flatnonzero(keep)
    upwind_m = sps.coo_matrix((np.ones(rows.size), (rows, up[rows])), shape=(nf, nc)).tocsr()
    sgn_div = np.asarray(sps.csr_matrix(g.cell_faces).sum(axis=1)).ravel()
    neu = np.flatnonzero(is_neu)
    inf = np.flatnonzero(inflow)
    return {"upwind": upwind_m,
            "bound_transport_neu": sps.coo_matrix((sgn_div[neu], (neu, neu)), shape=(nf, nf)).tocsr(),
            "bound_transport_dir": sps.coo_matrix((np.ones(inf.size), (inf, inf)), shape=(nf, nf)).tocsr()}
