"""CPU oracle for the MPFA / MPSA / Biot hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``porepy_b200``) never does
and fails loudly when its CUDA library is missing.

What it restates (reference = pmgbergen/porepy v1.11.0, /root/reference):

* sub-cell topology            src/porepy/numerics/fv/_fvutils.py:51-172
* continuity-point distances   _fvutils.py:222-277  (eta -> 0 on boundary faces, :259-263)
* n.K products                 _fvutils.py:697-762
* MPFA local systems + outputs src/porepy/numerics/fv/mpfa.py:592-1156,
                               boundary rhs :1414-1578, vector source :1158-1307,
                               pressure trace :1628-1690
* MPSA local systems + outputs src/porepy/numerics/fv/mpsa.py:531-930, n.C split
                               :1461-1675, ncasym elimination :1932-2000, bound rhs
                               :984-1185, displacement trace :1187-1275
* Biot coupling terms          src/porepy/numerics/fv/biot.py:714-1135
* block inversion              src/porepy/numerics/linalg/matrix_operations.py:1175-1371
                               (dense ``np.linalg.inv`` per interaction region)
* row scaling                  matrix_operations.py:1880-1906

Like the reference, every interaction region (node) gets ONE dense local
system over the sub-cell gradients (order nd*#subcells for MPFA, nd^2*#subcells
for MPSA), rows scaled by 1/sum|row|, inverted with LAPACK (``np.linalg.inv``)
and multiplied out into the face-indexed matrices.  Unlike the reference the
global sparse algebra is replaced by batched dense algebra over groups of nodes
with equal block size (pure restructuring; the arithmetic per node is the same).

Parity pin: validated against the reference itself in this container
(tools/make_golden.py -> tests/golden/*.npz, tests/test_oracle_vs_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sps

# ----------------------------------------------------------------------------------
# topology
# ----------------------------------------------------------------------------------


@dataclass
class SubcellTopology:
    """Sub-half-face enumeration, lexsorted by (cell, node, face)  (_fvutils.py:143-150).

    h-indexed arrays (one entry per (cell, face, node) incidence): cno, nno, fno,
    subfno, sgn.  Because every sub-cell has exactly nd sub-half-faces
    (_fvutils.py:735) sub-cell id = h // nd and gradient slot = h % nd.
    """

    nd: int
    nc: int
    nf: int
    nn: int
    cno: np.ndarray
    nno: np.ndarray
    fno: np.ndarray
    subfno: np.ndarray
    sgn: np.ndarray
    # per sub-face (id = position in face_nodes.indices)
    sf_node: np.ndarray
    sf_face: np.ndarray
    sf_h1: np.ndarray  # unique side = smaller cell index (_fvutils.py:163)
    sf_h2: np.ndarray  # second side or -1
    num_face_nodes: np.ndarray
    # per sub-cell
    sc_cell: np.ndarray
    sc_node: np.ndarray
    loc_sc: np.ndarray  # index of the sub-cell inside its node
    nsc_node: np.ndarray  # sub-cells per node
    loc_sf: np.ndarray  # index of the sub-face inside its node
    nsf_node: np.ndarray


def subcell_topology(g) -> SubcellTopology:
    nd = int(g.dim)
    cf = sps.csc_matrix(g.cell_faces)
    cf.sort_indices()
    fn = sps.csc_matrix(g.face_nodes)
    nf, nc = cf.shape
    nn = fn.shape[0]
    face_ind = cf.indices.astype(np.int64)
    cell_ind = np.repeat(np.arange(nc, dtype=np.int64), np.diff(cf.indptr))
    sgn_cf = np.asarray(cf.data, dtype=np.float64)
    nfn = np.diff(fn.indptr).astype(np.int64)
    reps = nfn[face_ind]
    tot = int(reps.sum())
    cells_dup = np.repeat(cell_ind, reps)
    faces_dup = np.repeat(face_ind, reps)
    sgn_dup = np.repeat(sgn_cf, reps)
    offs = np.arange(tot) - np.repeat(np.cumsum(reps) - reps, reps)
    pos = np.repeat(fn.indptr[face_ind].astype(np.int64), reps) + offs
    nodes_dup = fn.indices[pos].astype(np.int64)
    idx = np.lexsort((pos, faces_dup, nodes_dup, cells_dup))
    cno, nno, fno, subfno, sgn = (
        cells_dup[idx], nodes_dup[idx], faces_dup[idx], pos[idx], sgn_dup[idx])
    H = cno.size
    # every (cell, node) pair must own exactly nd sub-half-faces (_fvutils.py:735)
    if H % nd != 0:
        raise AssertionError("cells must have exactly nd faces meeting in each vertex")
    key = cno * nn + nno
    k2 = key.reshape(-1, nd)
    if not np.all(k2 == k2[:, :1]) or np.any(np.diff(k2[:, 0]) <= 0):
        raise AssertionError("cells must have exactly nd faces meeting in each vertex")
    sc_cell = cno[::nd].copy()
    sc_node = nno[::nd].copy()
    nsc_node = np.bincount(sc_node, minlength=nn)
    order = np.argsort(sc_node, kind="stable")
    ptr = np.cumsum(nsc_node) - nsc_node
    loc_sc = np.empty(sc_node.size, dtype=np.int64)
    loc_sc[order] = np.arange(sc_node.size) - ptr[sc_node[order]]

    U = fn.indices.size
    sf_node = fn.indices.astype(np.int64)
    sf_face = np.repeat(np.arange(nf, dtype=np.int64), nfn)
    horder = np.argsort(subfno, kind="stable")
    first = np.full(U, -1, dtype=np.int64)
    last = np.full(U, -1, dtype=np.int64)
    # stable sort: within equal subfno the h order (cell-major) is kept
    sfs = subfno[horder]
    starts = np.flatnonzero(np.r_[True, sfs[1:] != sfs[:-1]])
    ends = np.r_[starts[1:], sfs.size] - 1
    first[sfs[starts]] = horder[starts]
    last[sfs[ends]] = horder[ends]
    if np.any(first < 0):
        raise AssertionError("face without neighbouring cell")
    if np.any(ends - starts > 1):
        raise AssertionError("face with more than two neighbouring cells")
    sf_h2 = np.where(last != first, last, -1)
    nsf_node = np.bincount(sf_node, minlength=nn)
    order_f = np.argsort(sf_node, kind="stable")
    ptr_f = np.cumsum(nsf_node) - nsf_node
    loc_sf = np.empty(U, dtype=np.int64)
    loc_sf[order_f] = np.arange(U) - ptr_f[sf_node[order_f]]
    return SubcellTopology(nd, nc, nf, nn, cno, nno, fno, subfno, sgn, sf_node, sf_face,
                           first, sf_h2, nfn, sc_cell, sc_node, loc_sc, nsc_node,
                           loc_sf, nsf_node)


def continuity_dist(g, st: SubcellTopology, eta: float) -> np.ndarray:
    """d[h] = x_cp - x_cell (nd x H);  x_cp = x_f + eta (x_node - x_f), eta=0 on
    boundary faces (_fvutils.py:254-266)."""
    nd = st.nd
    bnd_face = np.bincount(st.fno, minlength=st.nf) == st.num_face_nodes  # single side
    eta_h = np.where(bnd_face[st.fno], 0.0, float(eta))
    fc = g.face_centers[:nd, st.fno]
    cp = fc + eta_h * (g.nodes[:nd, st.nno] - fc)
    return cp - g.cell_centers[:nd, st.cno]


def determine_eta(g) -> float:
    """_fvutils.py:280-305."""
    name = getattr(g, "name", "")
    if not isinstance(name, str):
        name = " ".join(str(n) for n in name)
    return 1.0 / 3.0 if ("TriangleGrid" in name or "TetrahedralGrid" in name) else 0.0


# ----------------------------------------------------------------------------------
# batched local-system machinery shared by MPFA and MPSA
# ----------------------------------------------------------------------------------


class _NodeGroups:
    """Groups of nodes with equal local-system order; dense batched scatter/inverse."""

    def __init__(self, n_of_node: np.ndarray, max_batch_elems: float = 4.0e7):
        self.n_of_node = n_of_node
        self.groups = []
        for n in np.unique(n_of_node):
            if n == 0:
                continue
            nodes = np.flatnonzero(n_of_node == n)
            chunk = max(1, int(max_batch_elems // (n * n)))
            for i in range(0, nodes.size, chunk):
                self.groups.append((int(n), nodes[i:i + chunk]))


def _ranges(ptr: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Concatenate arange(ptr[i], ptr[i+1]) for i in ids (vectorised)."""
    lens = ptr[ids + 1] - ptr[ids]
    tot = int(lens.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.repeat(ptr[ids] - (np.cumsum(lens) - lens), lens)
    return starts + np.arange(tot)


def _invert_scaled(A: np.ndarray):
    """Row scaling 1/sum|row| (matrix_operations.py:1880-1906, mpfa.py:1013-1016) then
    dense inverse per block (matrix_operations.py:1370).  Returns inv(A) (scaling folded
    back in, mpfa.py:1045)."""
    s = np.abs(A).sum(axis=2)
    if np.any(s == 0):
        raise ValueError("Error in inversion of local linear systems")
    scale = 1.0 / s
    try:
        inv = np.linalg.inv(A * scale[:, :, None])
    except np.linalg.LinAlgError as e:  # matrix_operations.py:1487-1490
        raise ValueError("Error in inversion of local linear systems") from e
    return inv * scale[:, None, :]


def _coo(rows, cols, vals, shape):
    m = sps.coo_matrix((vals, (rows, cols)), shape=shape).tocsr()
    m.sum_duplicates()
    return m


# ----------------------------------------------------------------------------------
# MPFA
# ----------------------------------------------------------------------------------


def _scalar_bc_flags(bc, st: SubcellTopology):
    """Per-face flags.  Internal (fracture) faces -> Neumann (mpfa.py:1452-1454)."""
    nf = st.nf
    bnd_face = np.bincount(st.fno, minlength=nf) == st.num_face_nodes
    internal = np.asarray(getattr(bc, "is_internal", np.zeros(nf, bool)), bool)
    is_dir = np.asarray(bc.is_dir, bool) & ~internal & bnd_face
    is_rob = np.asarray(bc.is_rob, bool) & ~internal & bnd_face
    is_neu = bnd_face & ~is_dir & ~is_rob
    rw = np.asarray(getattr(bc, "robin_weight", np.ones(nf)), dtype=float)
    return bnd_face, is_dir, is_neu, is_rob, rw


def mpfa(g, k_values: np.ndarray, bc, eta: float | None = None) -> dict:
    """MPFA-O discretization; returns the six matrices of mpfa.py:496-508.

    k_values: (3,3,nc) as ``SecondOrderTensor.values``.  bc: object with is_dir /
    is_neu / is_rob / is_internal / robin_weight (face arrays).
    Keys: flux, bound_flux, bound_pressure_cell, bound_pressure_face, vector_source,
    bound_pressure_vector_source.
    """
    st = subcell_topology(g)
    nd, nc, nf = st.nd, st.nc, st.nf
    if eta is None:
        eta = determine_eta(g)
    H = st.cno.size
    h = np.arange(H)
    sc_h = h // nd
    slot_h = h % nd
    bnd_face, is_dir, is_neu, is_rob, rw = _scalar_bc_flags(bc, st)
    m_f = st.num_face_nodes.astype(float)
    d = continuity_dist(g, st, eta)  # nd x H
    # r_h = (n_f / m_f)^T K_c          (_fvutils.py:697-762)
    nsub = g.face_normals[:nd, st.fno] / m_f[st.fno]
    Kc = k_values[:nd, :nd, :][:, :, st.cno]
    r = np.einsum("ih,ijh->jh", nsub, Kc)  # nd x H

    # ---- row numbering inside each node
    U = st.sf_node.size
    f_u = st.sf_face
    u_dir, u_neu, u_rob = is_dir[f_u], is_neu[f_u], is_rob[f_u]
    has_flux = ~u_dir  # interior, Neumann, Robin
    has_pres = ~(u_neu | u_rob)  # interior, Dirichlet
    order = np.argsort(st.sf_node, kind="stable")
    cnt = has_flux.astype(np.int64) + has_pres.astype(np.int64)
    cs = np.cumsum(cnt[order]) - cnt[order]
    node_start = np.zeros(st.nn + 1, dtype=np.int64)
    np.add.at(node_start, st.sf_node + 1, cnt)
    node_start = np.cumsum(node_start)
    row0 = np.empty(U, dtype=np.int64)
    row0[order] = cs - node_start[st.sf_node[order]]
    row_flux = np.where(has_flux, row0, -1)
    row_pres = np.where(has_pres, row0 + has_flux, -1)
    n_node = nd * st.nsc_node
    if not np.array_equal(node_start[1:] - node_start[:-1], n_node):
        raise AssertionError("local systems are not square")
    # boundary sub-faces: local column inside the node
    u_bnd = bnd_face[f_u]
    nb_node = np.bincount(st.sf_node[u_bnd], minlength=st.nn)
    ob = np.argsort(st.sf_node[u_bnd], kind="stable")
    ptr_b = np.cumsum(nb_node) - nb_node
    loc_b = np.full(U, -1, dtype=np.int64)
    ub_idx = np.flatnonzero(u_bnd)
    loc_b[ub_idx[ob]] = np.arange(ub_idx.size) - ptr_b[st.sf_node[ub_idx[ob]]]

    u_h = st.subfno  # sub-face of each sub-half-face
    node_h = st.nno
    col_h = st.loc_sc[sc_h] * nd  # first gradient column of the sub-cell of h
    area_sub = g.face_areas[st.fno] / m_f[st.fno]

    out_rows = {k: [] for k in ("flux", "bf", "bpc", "bpf", "vs", "bpvs")}
    out_cols = {k: [] for k in out_rows}
    out_vals = {k: [] for k in out_rows}

    # node -> list of its sub-cells / sub-faces (global ids), padded per group
    sc_order = np.argsort(st.sc_node, kind="stable")
    sc_ptr = np.r_[0, np.cumsum(st.nsc_node)]
    sf_order = order
    sf_ptr = np.r_[0, np.cumsum(st.nsf_node)]
    h_order = np.argsort(node_h, kind="stable")
    h_ptr = np.r_[0, np.cumsum(np.bincount(node_h, minlength=st.nn))]

    for n, nodes in _NodeGroups(n_node).groups:
        B = nodes.size
        nsc = n // nd
        bidx = np.full(st.nn, -1, dtype=np.int64)
        bidx[nodes] = np.arange(B)
        # sub-half-faces of these nodes
        hh = h_order[_ranges(h_ptr, nodes)]
        b_h = bidx[node_h[hh]]
        uu = u_h[hh]
        rf = row_flux[uu]
        rp = row_pres[uu]
        robh = u_rob[uu]
        A = np.zeros((B, n, n))
        Cc = np.zeros((B, n, nsc))
        V = np.zeros((B, n, n))
        for j in range(nd):
            cj = col_h[hh] + j
            mk = rf >= 0
            val = st.sgn[hh] * r[j, hh]
            A[b_h[mk], rf[mk], cj[mk]] += val[mk]
            V[b_h[mk], rf[mk], cj[mk]] += val[mk]
            # Robin: - w * a * d  (mpfa.py:869-887,997)
            mr = mk & robh
            A[b_h[mr], rf[mr], cj[mr]] -= (rw[st.fno[hh]] * area_sub[hh] * d[j, hh])[mr]
            mp = rp >= 0
            A[b_h[mp], rp[mp], cj[mp]] += (st.sgn[hh] * d[j, hh])[mp]
        mp = rp >= 0
        Cc[b_h[mp], rp[mp], st.loc_sc[sc_h[hh]][mp]] += st.sgn[hh][mp]
        mr = (rf >= 0) & robh
        Cc[b_h[mr], rf[mr], st.loc_sc[sc_h[hh]][mr]] -= (rw[st.fno[hh]] * area_sub[hh])[mr]
        # boundary rhs (mpfa.py:1414-1578)
        su = sf_order[_ranges(sf_ptr, nodes)]  # sub-faces of the group
        b_u = bidx[st.sf_node[su]]
        nb = int(nb_node[nodes].max()) if B else 0
        nsf = int(st.nsf_node[nodes].max())
        Bb = np.zeros((B, n, max(nb, 1)))
        ub = su[u_bnd[su]]
        if ub.size:
            bb = bidx[st.sf_node[ub]]
            fl = (u_neu | u_rob)[ub]
            Bb[bb[fl], row_flux[ub[fl]], loc_b[ub[fl]]] = -1.0 / m_f[f_u[ub[fl]]]
            di = u_dir[ub]
            Bb[bb[di], row_pres[ub[di]], loc_b[ub[di]]] = st.sgn[st.sf_h1[ub[di]]]
        Ainv = _invert_scaled(A)
        Gc = -Ainv @ Cc
        Gb = Ainv @ Bb
        Gv = Ainv @ V
        # output functionals per sub-face
        R1 = np.zeros((B, nsf, n))
        T = np.zeros((B, nsf, n))
        E = np.zeros((B, nsf, nsc))
        Rown = np.zeros((B, nsf, n))  # + r at own sub-cell (vector_source_faces)
        h1 = st.sf_h1[su]
        h2 = st.sf_h2[su]
        ls = st.loc_sf[su]
        nside = np.where(h2 >= 0, 2.0, 1.0)
        for j in range(nd):
            R1[b_u, ls, col_h[h1] + j] = r[j, h1]
            T[b_u, ls, col_h[h1] + j] += d[j, h1] / nside
            m2 = h2 >= 0
            T[b_u[m2], ls[m2], col_h[h2[m2]] + j] += d[j, h2[m2]] / nside[m2]
        E[b_u, ls, st.loc_sc[sc_h[h1]]] += 1.0 / nside
        m2 = h2 >= 0
        E[b_u[m2], ls[m2], st.loc_sc[sc_h[h2[m2]]]] += 1.0 / nside[m2]
        fl_c = -R1 @ Gc
        fl_b = -R1 @ Gb
        vs = -R1 @ Gv + R1
        inv_m = np.zeros((B, nsf, 1))
        inv_m[b_u, ls, 0] = 1.0 / m_f[f_u[su]]
        pc = (T @ Gc + E) * inv_m
        pb = (T @ Gb) * inv_m
        pv = (T @ Gv) * inv_m
        # global ids
        face_of = np.full((B, nsf), -1, dtype=np.int64)
        face_of[b_u, ls] = f_u[su]
        sc_g = sc_order[_ranges(sc_ptr, nodes)]
        cell_of = np.full((B, nsc), -1, dtype=np.int64)
        cell_of[bidx[st.sc_node[sc_g]], st.loc_sc[sc_g]] = st.sc_cell[sc_g]
        bface_of = np.full((B, max(nb, 1)), -1, dtype=np.int64)
        if ub.size:
            bface_of[bidx[st.sf_node[ub]], loc_b[ub]] = f_u[ub]
        vcol_of = (cell_of[:, :, None] * nd + np.arange(nd)[None, None, :]).reshape(B, n)
        vcol_of[np.repeat(cell_of, nd, axis=1) < 0] = -1

        def emit(key, M, colmap):
            rr = np.broadcast_to(face_of[:, :, None], M.shape)
            cc = np.broadcast_to(colmap[:, None, :], M.shape)
            ok = (rr >= 0) & (cc >= 0)
            out_rows[key].append(rr[ok])
            out_cols[key].append(cc[ok])
            out_vals[key].append(M[ok])

        emit("flux", fl_c, cell_of)
        emit("bf", fl_b, bface_of)
        emit("bpc", pc, cell_of)
        emit("bpf", pb, bface_of)
        emit("vs", vs, vcol_of)
        emit("bpvs", pv, vcol_of)

    def fin(key, shape):
        if not out_rows[key]:
            return sps.csr_matrix(shape)
        return _coo(np.concatenate(out_rows[key]), np.concatenate(out_cols[key]),
                    np.concatenate(out_vals[key]), shape)

    return {
        "flux": fin("flux", (nf, nc)),
        "bound_flux": fin("bf", (nf, nf)),
        "bound_pressure_cell": fin("bpc", (nf, nc)),
        "bound_pressure_face": fin("bpf", (nf, nf)),
        "vector_source": fin("vs", (nf, nc * nd)),
        "bound_pressure_vector_source": fin("bpvs", (nf, nc * nd)),
    }


# ----------------------------------------------------------------------------------
# MPSA (+ Biot coupling terms)
# ----------------------------------------------------------------------------------

_SYM_MASK_3D = np.eye(9, dtype=bool)
for _a, _b in ((0, 4), (0, 8), (4, 0), (4, 8), (8, 0), (8, 4)):
    _SYM_MASK_3D[_a, _b] = True
_SYM_MASK_2D = np.eye(4, dtype=bool)
_SYM_MASK_2D[0, 3] = _SYM_MASK_2D[3, 0] = True


def split_stiffness(c_values: np.ndarray, nd: int):
    """mpsa.py:1461-1518.  c_values (9,9,nc) -> csym, casym (nd^2, nd^2, nc)."""
    c = np.asarray(c_values, dtype=float)
    if nd == 2 and c.shape[0] == 9:
        keep = [0, 1, 3, 4]
        c = c[np.ix_(keep, keep)]
    mask = _SYM_MASK_3D if nd == 3 else _SYM_MASK_2D
    csym = np.where(mask[:, :, None], c, 0.0)
    return csym, c - csym


def _vector_bc_flags(bc, st: SubcellTopology):
    nd, nf = st.nd, st.nf
    bnd_face = np.bincount(st.fno, minlength=nf) == st.num_face_nodes
    is_dir = np.asarray(bc.is_dir, bool)[:nd] & bnd_face
    is_rob = np.asarray(bc.is_rob, bool)[:nd] & bnd_face
    is_neu = bnd_face[None, :] & ~is_dir & ~is_rob
    rw = getattr(bc, "robin_weight", None)
    if rw is None:
        rw = np.tile(np.eye(nd)[:, :, None], (1, 1, nf))
    rw = np.asarray(rw, dtype=float)[:nd, :nd]
    basis = getattr(bc, "basis", None)
    if basis is not None:
        b = np.asarray(basis, dtype=float)[:nd, :nd]
        if not np.allclose(b, np.eye(nd)[:, :, None]):
            raise NotImplementedError("oracle: rotated boundary bases are not restated")
    return bnd_face, is_dir, is_neu, is_rob, rw


def mpsa(g, c_values: np.ndarray, bc, eta: float | None = None, alpha: dict | None = None
         ) -> dict:
    """MPSA-W discretization (mpsa.py:531-781); with ``alpha`` ({key: (nd,nd,nc) or
    (3,3,nc) array}) also the Biot coupling terms of biot.py:714-878.

    Keys: stress, bound_stress, bound_displacement_cell, bound_displacement_face; per
    coupling key in ``alpha`` additionally (dict-valued, names as in biot.py:94-111)
    displacement_divergence, boundary_displacement_divergence, scalar_gradient,
    mpsa_consistency, bound_displacement_pressure.
    """
    st = subcell_topology(g)
    nd, nc, nf = st.nd, st.nc, st.nf
    nd2 = nd * nd
    if eta is None:
        eta = determine_eta(g)
    H = st.cno.size
    sc_h = np.arange(H) // nd
    bnd_face, is_dir, is_neu, is_rob, rw = _vector_bc_flags(bc, st)
    m_f = st.num_face_nodes.astype(float)
    d = continuity_dist(g, st, eta)
    nsub = g.face_normals[:nd, st.fno] / m_f[st.fno]
    csym, casym = split_stiffness(c_values, nd)
    # node-volume weights (mpsa.py:1619-1640)
    ncn = np.bincount(st.sc_cell, minlength=nc).astype(float)
    cvol = g.cell_volumes / ncn
    node_vol = np.bincount(st.sc_node, weights=cvol[st.sc_cell], minlength=st.nn)
    w_sc = cvol[st.sc_cell] / node_vol[st.sc_node]

    U = st.sf_node.size
    f_u = st.sf_face
    u_dir, u_neu, u_rob = is_dir[:, f_u], is_neu[:, f_u], is_rob[:, f_u]  # nd x U
    has_str = ~u_dir
    has_dis = ~(u_neu | u_rob)
    cnt = (has_str.astype(np.int64) + has_dis.astype(np.int64))  # nd x U
    cnt_u = cnt.sum(axis=0)
    order = np.argsort(st.sf_node, kind="stable")
    cs = np.cumsum(cnt_u[order]) - cnt_u[order]
    node_start = np.zeros(st.nn + 1, dtype=np.int64)
    np.add.at(node_start, st.sf_node + 1, cnt_u)
    node_start = np.cumsum(node_start)
    row0 = np.empty(U, dtype=np.int64)
    row0[order] = cs - node_start[st.sf_node[order]]
    within = np.cumsum(cnt, axis=0) - cnt  # rows of components before i
    row_str = np.where(has_str, row0[None, :] + within, -1)
    row_dis = np.where(has_dis, row0[None, :] + within + has_str, -1)
    n_node = nd2 * st.nsc_node
    if not np.array_equal(node_start[1:] - node_start[:-1], n_node):
        raise AssertionError("local systems are not square")
    u_bnd = bnd_face[f_u]
    nb_node = np.bincount(st.sf_node[u_bnd], minlength=st.nn)
    ub_idx = np.flatnonzero(u_bnd)
    ob = np.argsort(st.sf_node[ub_idx], kind="stable")
    ptr_b = np.cumsum(nb_node) - nb_node
    loc_b = np.full(U, -1, dtype=np.int64)
    loc_b[ub_idx[ob]] = np.arange(ub_idx.size) - ptr_b[st.sf_node[ub_idx[ob]]]
    # _eliminate_ncasym (mpsa.py:1932-2000): per node and component
    elim = np.zeros((nd, st.nn), dtype=bool)
    elim_r = np.zeros((nd, st.nn), dtype=bool)
    for i in range(nd):
        elim[i] = st.nsc_node < np.bincount(st.sf_node[u_neu[i]], minlength=st.nn)
        elim_r[i] = st.nsc_node < np.bincount(st.sf_node[u_rob[i]], minlength=st.nn)

    u_h = st.subfno
    node_h = st.nno
    col_h = st.loc_sc[sc_h] * nd2
    area_sub = g.face_areas[st.fno] / m_f[st.fno]

    keys = ["stress", "bs", "bdc", "bdf"]
    alpha = alpha or {}
    al = {}
    for ak, av in alpha.items():
        av = np.asarray(av, dtype=float)
        if av.ndim == 0 or av.ndim == 1:
            av = np.eye(nd)[:, :, None] * np.broadcast_to(av, (nc,))[None, None, :]
        al[ak] = av[:nd, :nd]
        keys += [f"{x}:{ak}" for x in ("dd", "bdd", "sg", "cons", "bdp")]
    out = {k: ([], [], []) for k in keys}

    sc_order = np.argsort(st.sc_node, kind="stable")
    sc_ptr = np.r_[0, np.cumsum(st.nsc_node)]
    sf_ptr = np.r_[0, np.cumsum(st.nsf_node)]
    h_order = np.argsort(node_h, kind="stable")
    h_ptr = np.r_[0, np.cumsum(np.bincount(node_h, minlength=st.nn))]

    for n, nodes in _NodeGroups(n_node, max_batch_elems=2.0e7).groups:
        B = nodes.size
        nsc = n // nd2
        bidx = np.full(st.nn, -1, dtype=np.int64)
        bidx[nodes] = np.arange(B)
        hh = h_order[_ranges(h_ptr, nodes)]
        b_h = bidx[node_h[hh]]
        uu = u_h[hh]
        sg = st.sgn[hh]
        lsc_h = st.loc_sc[sc_h[hh]]
        su = order[_ranges(sf_ptr, nodes)]
        b_u = bidx[st.sf_node[su]]
        ls = st.loc_sf[su]
        nsf = int(st.nsf_node[nodes].max())
        nb = max(int(nb_node[nodes].max()), 1)
        sc_g = sc_order[_ranges(sc_ptr, nodes)]
        b_sc = bidx[st.sc_node[sc_g]]
        l_sc = st.loc_sc[sc_g]
        # weighted asymmetric tensors of the node's sub-cells: (B, nsc, nd2, nd2)
        CAw = np.zeros((B, nsc, nd2, nd2))
        CAw[b_sc, l_sc] = np.moveaxis(casym[:, :, st.sc_cell[sc_g]], 2, 0) * w_sc[sc_g][:, None, None]
        CAw_rows = CAw.transpose(0, 2, 1, 3).reshape(B, nd2, nsc * nd2)  # [b, p, (k', q)]

        A = np.zeros((B, n, n))
        Cc = np.zeros((B, n, nsc * nd))
        Bb = np.zeros((B, n, nb * nd))
        Pj = {ak: np.zeros((B, n, nsc)) for ak in al}
        cs_h = csym[:, :, st.cno[hh]]  # nd2 x nd2 x Hg
        for i in range(nd):
            rs = row_str[i, uu]
            rd = row_dis[i, uu]
            ms = rs >= 0
            # symmetric part, own sub-cell:  sum_r n_r C[(i,r), q]
            tsym = np.einsum("rh,rqh->qh", nsub[:, hh], cs_h[i * nd:(i + 1) * nd])  # nd2 x Hg
            for q in range(nd2):
                A[b_h[ms], rs[ms], col_h[hh][ms] + q] += (sg * tsym[q])[ms]
            # asymmetric part on Neumann / Robin boundary rows (single side)
            bn = ms & (u_neu[i, uu] | u_rob[i, uu])
            keep = bn & ~np.where(u_neu[i, uu], elim[i, node_h[hh]], elim_r[i, node_h[hh]])
            if np.any(keep):
                hk = np.flatnonzero(keep)
                tas = np.einsum("rh,hrc->hc", nsub[:, hh[hk]],
                                CAw_rows[b_h[hk], i * nd:(i + 1) * nd, :])
                A[b_h[hk], rs[hk], :] += sg[hk, None] * tas
            # Robin: + a * sum_j w[i,j] (d.G_j + u_j)  (mpsa.py:1381-1459)
            mr = ms & u_rob[i, uu]
            if np.any(mr):
                hr = np.flatnonzero(mr)
                wa = rw[i][:, st.fno[hh[hr]]] * area_sub[hh[hr]]  # nd x nr
                for j in range(nd):
                    for kk in range(nd):
                        A[b_h[hr], rs[hr], col_h[hh[hr]] + j * nd + kk] += wa[j] * d[kk, hh[hr]]
                    Cc[b_h[hr], rs[hr], lsc_h[hr] * nd + j] += wa[j]
            md = rd >= 0
            for kk in range(nd):
                A[b_h[md], rd[md], col_h[hh][md] + i * nd + kk] += (sg * d[kk, hh])[md]
            Cc[b_h[md], rd[md], lsc_h[md] * nd + i] += sg[md]
            # Biot pressure-jump rhs on the stress rows (biot.py:969-1019)
            for ak, av in al.items():
                na = np.einsum("rh,rh->h", nsub[:, hh], av[:, i, st.cno[hh]])  # (n^T alpha)_i
                Pj[ak][b_h[ms], rs[ms], lsc_h[ms]] += (sg * na)[ms]
            # boundary rhs (mpsa.py:984-1185)
            ub = su[u_bnd[su]]
            if ub.size:
                bb = bidx[st.sf_node[ub]]
                fl = (u_neu | u_rob)[i, ub]
                Bb[bb[fl], row_str[i, ub[fl]], loc_b[ub[fl]] * nd + i] = 1.0 / m_f[f_u[ub[fl]]]
                di = u_dir[i, ub]
                Bb[bb[di], row_dis[i, ub[di]], loc_b[ub[di]] * nd + i] = st.sgn[st.sf_h1[ub[di]]]
        Ainv = _invert_scaled(A)
        Gc = -Ainv @ Cc
        Gb = Ainv @ Bb
        # hook (unique side) and trace functionals, rows (ls*nd + i)
        h1 = st.sf_h1[su]
        h2 = st.sf_h2[su]
        m2 = h2 >= 0
        nside = np.where(m2, 2.0, 1.0)
        Hk = np.zeros((B, nsf * nd, n))
        T = np.zeros((B, nsf * nd, n))
        E = np.zeros((B, nsf * nd, nsc * nd))
        cs_1 = csym[:, :, st.cno[h1]]
        bndu = u_bnd[su]
        for i in range(nd):
            tsym = np.einsum("rh,rqh->qh", nsub[:, h1], cs_1[i * nd:(i + 1) * nd])
            for q in range(nd2):
                Hk[b_u, ls * nd + i, col_h[h1] + q] += tsym[q]
            tas = np.einsum("rh,hrc->hc", nsub[:, h1], CAw_rows[b_u, i * nd:(i + 1) * nd, :])
            zero = bndu & ((u_neu[i, su] & elim[i, st.sf_node[su]])
                           | (u_rob[i, su] & elim_r[i, st.sf_node[su]]))
            tas[zero] = 0.0
            Hk[b_u, ls * nd + i, :] += tas
            for kk in range(nd):
                T[b_u, ls * nd + i, col_h[h1] + i * nd + kk] += d[kk, h1] / nside
                T[b_u[m2], ls[m2] * nd + i, col_h[h2[m2]] + i * nd + kk] += d[kk, h2[m2]] / nside[m2]
            E[b_u, ls * nd + i, st.loc_sc[sc_h[h1]] * nd + i] += 1.0 / nside
            E[b_u[m2], ls[m2] * nd + i, st.loc_sc[sc_h[h2[m2]]] * nd + i] += 1.0 / nside[m2]
        inv_m = np.zeros((B, nsf * nd, 1))
        for i in range(nd):
            inv_m[b_u, ls * nd + i, 0] = 1.0 / m_f[f_u[su]]
        # global ids
        frow = np.full((B, nsf * nd), -1, dtype=np.int64)
        for i in range(nd):
            frow[b_u, ls * nd + i] = f_u[su] * nd + i
        cell_of = np.full((B, nsc), -1, dtype=np.int64)
        cell_of[b_sc, l_sc] = st.sc_cell[sc_g]
        ccol = (cell_of[:, :, None] * nd + np.arange(nd)).reshape(B, nsc * nd)
        ccol[np.repeat(cell_of, nd, axis=1) < 0] = -1
        bface = np.full((B, nb), -1, dtype=np.int64)
        ub = su[u_bnd[su]]
        if ub.size:
            bface[bidx[st.sf_node[ub]], loc_b[ub]] = f_u[ub]
        bcol = (bface[:, :, None] * nd + np.arange(nd)).reshape(B, nb * nd)
        bcol[np.repeat(bface, nd, axis=1) < 0] = -1

        def emit(key, M, rowmap, colmap):
            rr = np.broadcast_to(rowmap[:, :, None], M.shape)
            cc = np.broadcast_to(colmap[:, None, :], M.shape)
            ok = (rr >= 0) & (cc >= 0)
            out[key][0].append(rr[ok])
            out[key][1].append(cc[ok])
            out[key][2].append(M[ok])

        emit("stress", Hk @ Gc, frow, ccol)
        emit("bs", Hk @ Gb, frow, bcol)
        emit("bdc", (T @ Gc + E) * inv_m, frow, ccol)
        emit("bdf", (T @ Gb) * inv_m, frow, bcol)
        for ak, av in al.items():
            Gp = Ainv @ Pj[ak]
            # dv_K = V_K/#nodes(K) * vec(alpha_K) at the sub-cell's slots (biot.py:1054-1135)
            DV = np.zeros((B, nsc, n))
            a_sc = av[:, :, st.sc_cell[sc_g]].reshape(nd2, -1)  # q = a*nd + k
            for q in range(nd2):
                DV[b_sc, l_sc, l_sc * nd2 + q] = cvol[st.sc_cell[sc_g]] * a_sc[q]
            emit(f"dd:{ak}", DV @ Gc, cell_of, ccol)
            emit(f"bdd:{ak}", DV @ Gb, cell_of, bcol)
            emit(f"cons:{ak}", DV @ Gp, cell_of, cell_of)
            sgm = Hk @ Gp
            # face term  -(n^T alpha_Ku)_i [K' == K_u]   (biot.py:853-855,1021-1036)
            for i in range(nd):
                na = np.einsum("rh,rh->h", nsub[:, h1], av[:, i, st.cno[h1]])
                sgm[b_u, ls * nd + i, st.loc_sc[sc_h[h1]]] -= na
            emit(f"sg:{ak}", sgm, frow, cell_of)
            emit(f"bdp:{ak}", (T @ Gp) * inv_m, frow, cell_of)

    def fin(key, shape):
        r_, c_, v_ = out[key]
        if not r_:
            return sps.csr_matrix(shape)
        return _coo(np.concatenate(r_), np.concatenate(c_), np.concatenate(v_), shape)

    res = {
        "stress": fin("stress", (nf * nd, nc * nd)),
        "bound_stress": fin("bs", (nf * nd, nf * nd)),
        "bound_displacement_cell": fin("bdc", (nf * nd, nc * nd)),
        "bound_displacement_face": fin("bdf", (nf * nd, nf * nd)),
    }
    if al:
        res["displacement_divergence"] = {ak: fin(f"dd:{ak}", (nc, nc * nd)) for ak in al}
        res["boundary_displacement_divergence"] = {ak: fin(f"bdd:{ak}", (nc, nf * nd)) for ak in al}
        res["scalar_gradient"] = {ak: fin(f"sg:{ak}", (nf * nd, nc)) for ak in al}
        res["mpsa_consistency"] = {ak: fin(f"cons:{ak}", (nc, nc)) for ak in al}
        res["bound_displacement_pressure"] = {ak: fin(f"bdp:{ak}", (nf * nd, nc)) for ak in al}
    return res
