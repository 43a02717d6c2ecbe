"""``DifferentiableTpfa`` -- the helper behind the reference's differentiable two-point flux (a permeability that
depends on the solution: ``DarcysLawAd`` / ``FouriersLawAd``, reference src/porepy/models/constitutive_laws.py:
1500-1583), reference src/porepy/numerics/fv/tpfa.py:281-760.

Two layers:

* the reference's grid-to-matrix helpers under their own names and conventions (``half_face_map``,
  ``half_face_geometry_matrices``, ``face_pairing_from_cell_array``, ``boundary_sign``, ``nd_to_3d`` and the
  filters), host-side scipy builders that a model's AD operator tree consumes once per grid;
* the evaluation the reference spells as an AD expression and re-parses in EVERY Newton iteration,

      t_hf = (d_vec @ n @ k_c) / dist,      T_f = 1 / (hf_to_f @ (1 / t_hf)),      dT_f / dk_c,

  as ONE kernel (``pb_tpfa_diff``, csrc/tpfa_diff.cuh ``tpfa_diff_face``: one thread per face): the face
  transmissibilities and their exact derivative with respect to the 9 * nc permeability entries (9 entries per
  half-face), returned as a value vector and a sparse Jacobian factor -- what ``AdArray`` carries as ``val`` / ``jac``
  after the chain ``one / (SparseArray(hf_to_f) @ (one / (SparseArray(d_n_by_dist) @ k_c)))``.

Half-faces are the non-zeros of ``cell_faces`` in the order ``scipy.sparse.find`` returns them (by face, then by
cell), as in the reference.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


def _find(cell_faces):
    """(face, cell, sign) of the half-faces in the reference's order (``sps.find``: sorted by face, then by cell)."""
    m = sps.csr_matrix(cell_faces, copy=True)   # CSC -> CSR: a counting sort by face; ascending cells inside a face
    m.eliminate_zeros()                          # sps.find drops explicit zeros
    m.sort_indices()
    fi = np.repeat(np.arange(m.shape[0], dtype=np.int64), np.diff(m.indptr))
    return fi, m.indices.astype(np.int64), np.asarray(m.data, dtype=np.float64)


def _expand(ind: np.ndarray, dim: int) -> np.ndarray:
    """pp.array_operations.expand_indices_nd: [i] -> [i*dim, ..., i*dim + dim - 1]."""
    return (np.repeat(np.asarray(ind, dtype=np.int64) * dim, dim) + np.tile(np.arange(dim), np.asarray(ind).size))


def _block_diag(blocks):
    if len(blocks) == 0:
        return sps.csr_matrix((0, 0))
    return sps.block_diag(blocks, format="csr") if len(blocks) > 1 else sps.csr_matrix(blocks[0])


class DifferentiableTpfa:
    """Mirror of ``pp.numerics.fv.tpfa.DifferentiableTpfa`` (tpfa.py:281); see the module docstring."""

    # ---- filters (tpfa.py:318-369): plain arrays here, the AD wrapping is the caller's
    def internal_boundary_filter(self, subdomains) -> np.ndarray:
        parts = [np.asarray(sd.tags["fracture_faces"]) for sd in subdomains]
        return np.hstack(parts) if parts else np.array([], dtype=int)

    def tip_filter(self, subdomains) -> np.ndarray:
        parts = [np.logical_and(sd.tags["tip_faces"], np.logical_not(sd.tags["domain_boundary_faces"]))
                 for sd in subdomains]
        return np.hstack(parts) if parts else np.array([], dtype=int)

    # ---- maps between cells, faces and half-faces (tpfa.py:401-481)
    def half_face_map(self, subdomains, from_entity: str = "half_faces", to_entity: str = "half_faces",
                      dimensions: tuple = (1, 1), with_sign: bool = False) -> sps.csr_matrix:
        def one(sd):
            fi, ci, sgn = _find(sd.cell_faces)
            indices, sizes = [], []
            for name in (to_entity, from_entity):
                if name == "cells":
                    indices.append(ci), sizes.append(sd.num_cells)
                elif name == "faces":
                    indices.append(fi), sizes.append(sd.num_faces)
                elif name == "half_faces":
                    indices.append(np.arange(fi.size)), sizes.append(fi.size)
                else:
                    raise ValueError(f"Unknown entity {name}.")
            rep_r = int(np.ceil(dimensions[1] / dimensions[0]))
            rep_c = int(np.ceil(dimensions[0] / dimensions[1]))
            assert dimensions[0] % dimensions[1] == 0 or dimensions[1] % dimensions[0] == 0
            rows = _expand(np.repeat(indices[0], rep_r), dimensions[0])
            cols = _expand(np.repeat(indices[1], rep_c), dimensions[1])
            vals = np.repeat(sgn, max(dimensions)) if with_sign else np.ones(cols.size)
            return sps.csr_matrix((vals, (rows, cols)), shape=(sizes[0] * dimensions[0], sizes[1] * dimensions[1]))
        return _block_diag([one(sd) for sd in subdomains])

    # ---- geometry per half-face (tpfa.py:483-660)
    def _cell_face_vectors(self, subdomains) -> sps.csr_matrix:
        def one(sd):
            fi, ci, _ = _find(sd.cell_faces)
            nhf = fi.size
            fc_cc = np.asarray(sd.face_centers)[:, fi] - np.asarray(sd.cell_centers)[:, ci]
            return sps.csr_matrix((fc_cc.ravel("F"), (np.repeat(np.arange(nhf), 3), _expand(np.arange(nhf), 3))),
                                  shape=(nhf, nhf * 3))
        return _block_diag([one(sd) for sd in subdomains])

    def _normal_vectors(self, subdomains) -> sps.csr_matrix:
        def one(sd):
            fi, ci, _ = _find(sd.cell_faces)
            nhf = fi.size
            n = np.asarray(sd.face_normals)
            rows = np.repeat(np.arange(nhf * 3), 3)
            cols = _expand(ci, 9)
            vals = n[:, np.repeat(fi, 3)].ravel("F")
            return sps.csr_matrix((vals, (rows, cols)), shape=(nhf * 3, sd.num_cells * 9))
        return _block_diag([one(sd) for sd in subdomains])

    def _cell_face_distances(self, subdomains) -> np.ndarray:
        vals = []
        for sd in subdomains:
            fi, ci, _ = _find(sd.cell_faces)
            fc_cc = np.asarray(sd.face_centers)[:, fi] - np.asarray(sd.cell_centers)[:, ci]
            vals.append(np.power(fc_cc, 2).sum(axis=0))
        return np.hstack(vals) if vals else np.array([])

    def half_face_geometry_matrices(self, subdomains):
        """(n, d_vec, dist): ``t_hf = d_vec @ n @ k_hf / dist`` (tpfa.py:617-660)."""
        return self._normal_vectors(subdomains), self._cell_face_vectors(subdomains), self._cell_face_distances(subdomains)

    def face_pairing_from_cell_array(self, subdomains) -> sps.csr_matrix:
        c_to_hf = self.half_face_map(subdomains, to_entity="half_faces", from_entity="cells")
        hf_to_f = self.half_face_map(subdomains, to_entity="faces", with_sign=True)
        return (hf_to_f @ c_to_hf).tocsr()

    def boundary_sign(self, subdomains) -> np.ndarray:
        out = []
        for sd in subdomains:
            fi, _, sgn = _find(sd.cell_faces)
            _, first = np.unique(fi, return_index=True)
            s = sgn[first].copy()
            is_int = np.logical_not(np.logical_or(sd.tags["domain_boundary_faces"], sd.tags["fracture_faces"]))
            s[is_int] = 0
            out.append(s)
        return np.hstack(out) if out else np.array([])

    def nd_to_3d(self, subdomains, nd: int, entity: str = "cells") -> sps.csr_matrix:
        def one(g):
            num = getattr(g, f"num_{entity}")
            rows = np.concatenate([np.arange(i, num * 3, 3) for i in range(nd)])
            cols = np.concatenate([np.arange(i, num * nd, nd) for i in range(nd)])
            return sps.csr_matrix((np.ones(cols.size), (rows, cols)), shape=(num * 3, num * nd))
        return _block_diag([one(g) for g in subdomains])

    # ---- the fused evaluation on the device
    def transmissibility(self, sd, k_c: np.ndarray, k_jac=None):
        """Face transmissibilities of ``sd`` for the cell-wise permeability ``k_c`` (9 * nc values, the 3 x 3 tensor
        of cell c at ``k_c[9c : 9c + 9]`` row-major: the reference's ``volumes * diffusivity_tensor`` AD vector,
        constitutive_laws.py:1544-1549).

        Returns ``(T_f, dT_dk, t_hf)``: the nf transmissibilities (the harmonic combination of the half-face values
        ``t_hf``), the Jacobian factor dT_f/dk_c as CSR (nf x 9 nc, 9 entries per half-face), and the half-face
        transmissibilities.  ``k_jac`` (the Jacobian dk_c/dx of an AD permeability, scipy sparse): the second return
        value is then dT_f/dx = dT_dk @ k_jac."""
        from . import fv
        fg = fv.FaceGrid.for_grid(sd)
        fi, ci, _ = _find(sd.cell_faces)
        nf, nc = sd.num_faces, sd.num_cells
        fc_ip = np.zeros(nf + 1, dtype=np.int32)
        np.cumsum(np.bincount(fi, minlength=nf), out=fc_ip[1:])
        k = np.ascontiguousarray(k_c, dtype=np.float64).reshape(-1)
        if k.size != 9 * nc:
            raise ValueError("k_c must hold 9 values per cell")
        t_hf, T, dT = fg.tpfa_diff(k, fc_ip)
        jac = sps.csr_matrix((dT, _expand(ci, 9), fc_ip.astype(np.int64) * 9), shape=(nf, 9 * nc))
        if k_jac is not None:
            jac = (jac @ sps.csr_matrix(k_jac)).tocsr()
        return T, jac, t_hf
