"""ctypes binding of tests/emu/_emu.so -- the host build of the CUDA node routines with a
1-thread team.  TEST INFRASTRUCTURE ONLY (see tests/emu/emu.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cpp")
LIB = os.path.join(HERE, "emu", "_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "porepy_b200", "csrc")


def _build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class EmuPlan:
    def __init__(self, g):
        L = lib()
        cf = sps.csc_matrix(g.cell_faces)
        fn = sps.csc_matrix(g.face_nodes)
        self.nd, self.nc, self.nf, self.nn = g.dim, g.num_cells, g.num_faces, g.num_nodes
        self._keep = [cf.indptr.astype(np.int32), cf.indices.astype(np.int32),
                      np.asarray(cf.data).astype(np.int8), fn.indptr.astype(np.int32),
                      fn.indices.astype(np.int32)]
        h = C.c_void_p()
        k = self._keep
        rc = L.emu_create(C.c_int(g.dim), C.c_int64(self.nc), C.c_int64(self.nf), C.c_int64(self.nn),
                          _p(k[0], C.c_int32), _p(k[1], C.c_int32), _p(k[2], C.c_int8),
                          _p(k[3], C.c_int32), _p(k[4], C.c_int32), C.byref(h))
        if rc == 3:
            raise AssertionError("cells must have exactly nd faces meeting in each vertex")
        if rc:
            raise RuntimeError(f"emu_create failed {rc}")
        self.h = h
        self.g = g
        self.geo = [np.ascontiguousarray(a, dtype=np.float64) for a in
                    (g.nodes, g.face_normals, g.face_centers, g.face_areas, g.cell_centers,
                     g.cell_volumes)]
        self.pat = [self._pattern(w) for w in range(4)]

    def _pattern(self, which):
        L = lib()
        nr, nz = C.c_int64(), C.c_int64()
        L.emu_pattern_size(self.h, C.c_int(which), C.byref(nr), C.byref(nz))
        ip = np.zeros(nr.value + 1, np.int32)
        ix = np.zeros(max(nz.value, 1), np.int32)
        L.emu_pattern_get(self.h, C.c_int(which), _p(ip, C.c_int32), _p(ix, C.c_int32))
        return ip, ix[:nz.value]

    def __del__(self):
        try:
            lib().emu_destroy(self.h)
        except Exception:
            pass

    # ---- expansions of the base patterns (same rules as porepy_b200/fv.py)
    def scalar(self, which, data, ncols):
        ip, ix = self.pat[which]
        return sps.csr_matrix((data, ix, ip), shape=(ip.size - 1, ncols))

    def cols_expanded(self, which, data, ncols, nd):
        ip, ix = self.pat[which]
        idx = (ix[:, None].astype(np.int64) * nd + np.arange(nd)).ravel()
        return sps.csr_matrix((data, idx, ip.astype(np.int64) * nd), shape=(ip.size - 1, ncols * nd))

    def mpfa(self, perm, bc_codes, robw, eta):
        L = lib()
        nd = self.nd
        nfc = self.pat[0][1].size
        nfb = self.pat[1][1].size
        out = [np.zeros(n) for n in (nfc, nfb, nfc, nfb, nfc * nd, nfc * nd)]
        perm = np.ascontiguousarray(perm, np.float64)
        bc_codes = np.ascontiguousarray(bc_codes, np.uint8)
        robw = None if robw is None else np.ascontiguousarray(robw, np.float64)
        rc = L.emu_mpfa(self.h, *[_p(a, C.c_double) for a in self.geo], _p(perm, C.c_double),
                        _p(bc_codes, C.c_uint8), _p(robw, C.c_double), C.c_double(eta),
                        *[_p(a, C.c_double) for a in out])
        if rc == 2:
            raise ValueError("Error in inversion of local linear systems")
        nc, nf = self.nc, self.nf
        return {
            "flux": self.scalar(0, out[0], nc),
            "bound_flux": self.scalar(1, out[1], nf),
            "bound_pressure_cell": self.scalar(0, out[2], nc),
            "bound_pressure_face": self.scalar(1, out[3], nf),
            "vector_source": self.cols_expanded(0, out[4], nc, nd),
            "bound_pressure_vector_source": self.cols_expanded(0, out[5], nc, nd),
        }


def block_expand(ip, ix, nd_rows, nd_cols):
    """indptr/indices of the matrix whose (r, c) base entries become nd_rows x nd_cols blocks
    laid out row r*nd_rows+i, column c*nd_cols+j; data position of (i, p, j) is
    nd_rows*nd_cols*ip[r] + i*nd_cols*len_r + (p-ip[r])*nd_cols + j."""
    ip = ip.astype(np.int64)
    lens = np.diff(ip)
    new_lens = np.repeat(lens * nd_cols, nd_rows)
    new_ip = np.r_[0, np.cumsum(new_lens)]
    # indices: for each base row r the expanded column list, repeated nd_rows times
    cols = (ix[:, None].astype(np.int64) * nd_cols + np.arange(nd_cols)).reshape(-1)
    # split per row and tile
    out = np.empty(new_ip[-1], dtype=np.int64)
    row_of = np.repeat(np.arange(lens.size), lens * nd_cols)  # row of each expanded col
    off_in_row = np.arange(cols.size) - np.repeat(ip[:-1] * nd_cols, lens * nd_cols)
    for i in range(nd_rows):
        dest = new_ip[row_of * nd_rows + i] + off_in_row
        out[dest] = cols
    return new_ip, out


def _mpsa(self, stiff, bc_codes, robw, eta, alpha=None, basis=None):
    L = lib()
    nd = self.nd
    nd2 = nd * nd
    nfc, nfb, ncc, ncb = (self.pat[w][1].size for w in range(4))
    out = [np.zeros(n) for n in (nfc * nd2, nfb * nd2, nfc * nd2, nfb * nd2)]
    alpha = alpha or {}
    keys = list(alpha)
    nal = len(keys)
    al = np.zeros((max(nal, 1), 3, 3, self.nc))
    for q, k in enumerate(keys):
        a = np.asarray(alpha[k], float)
        if a.ndim < 3:
            a = np.eye(3)[:, :, None] * np.broadcast_to(a, (self.nc,))
        al[q] = a
    bi = [[np.zeros(n) for _ in range(nal)] for n in (ncc * nd, ncb * nd, nfc * nd, ncc, nfc * nd)]
    PP = C.POINTER(C.c_double)
    arrs = [(PP * max(nal, 1))(*[_p(a, C.c_double) for a in lst]) if nal else (PP * 1)() for lst in bi]
    stiff = np.ascontiguousarray(stiff, np.float64)
    bc_codes = np.ascontiguousarray(bc_codes, np.uint8)
    robw = None if robw is None else np.ascontiguousarray(robw, np.float64)
    basis = None if basis is None else np.ascontiguousarray(np.asarray(basis, np.float64)[:nd, :nd])
    rc = L.emu_mpsa(self.h, *[_p(a, C.c_double) for a in self.geo], _p(stiff, C.c_double),
                    _p(bc_codes, C.c_uint8), _p(robw, C.c_double), _p(basis, C.c_double), C.c_double(eta),
                    C.c_int(nal),
                    _p(al, C.c_double), *[_p(a, C.c_double) for a in out], *arrs)
    if rc == 2:
        raise ValueError("Error in inversion of local linear systems")
    nc, nf = self.nc, self.nf

    def blk(which, data, nrows, ncols, br, bcn):
        ip, ix = block_expand(*self.pat[which], br, bcn)
        return sps.csr_matrix((data, ix, ip), shape=(nrows * br, ncols * bcn))

    res = {
        "stress": blk(0, out[0], nf, nc, nd, nd),
        "bound_stress": blk(1, out[1], nf, nf, nd, nd),
        "bound_displacement_cell": blk(0, out[2], nf, nc, nd, nd),
        "bound_displacement_face": blk(1, out[3], nf, nf, nd, nd),
    }
    if nal:
        res["displacement_divergence"] = {k: blk(2, bi[0][q], nc, nc, 1, nd) for q, k in enumerate(keys)}
        res["boundary_displacement_divergence"] = {k: blk(3, bi[1][q], nc, nf, 1, nd) for q, k in enumerate(keys)}
        res["scalar_gradient"] = {k: blk(0, bi[2][q], nf, nc, nd, 1) for q, k in enumerate(keys)}
        res["mpsa_consistency"] = {k: blk(2, bi[3][q], nc, nc, 1, 1) for q, k in enumerate(keys)}
        res["bound_displacement_pressure"] = {k: blk(0, bi[4][q], nf, nc, nd, 1) for q, k in enumerate(keys)}
    return res


EmuPlan.mpsa = _mpsa



class EmuBackedPlan:
    """Stand-in for ``porepy_b200.fv.DevicePlan`` with the calls the operator classes make, backed by
    the host build of the node routines.  TEST INFRASTRUCTURE: lets the CPU tests drive
    ``fv.Mpfa / Mpsa / Biot.discretize`` end to end (host logic + the kernels' source)."""

    def __init__(self, sd):
        from types import SimpleNamespace
        from porepy_b200 import fv
        arrs, self.rotation = fv.plan_geometry(sd)
        proxy = SimpleNamespace(dim=sd.dim, num_cells=sd.num_cells, num_faces=sd.num_faces,
                                num_nodes=sd.num_nodes, cell_faces=sd.cell_faces, face_nodes=sd.face_nodes,
                                nodes=arrs[0], face_normals=arrs[1], face_centers=arrs[2],
                                face_areas=arrs[3], cell_centers=arrs[4], cell_volumes=arrs[5])
        self.emu = EmuPlan(proxy)
        self.nd, self.nc = int(sd.dim), sd.num_cells

    @classmethod
    def for_grid(cls, sd):
        return cls(sd)

    def base_pattern(self, which):
        return self.emu.pat[which]

    def mpfa_upload(self, perm, codes, robw, eta):
        self._mpfa = (perm, codes, robw, eta)

    def mpfa_assemble(self, *a):
        self._out = self.emu.mpfa(*self._mpfa)
        return 0.0

    def mpfa_download(self, *a):
        return self._out

    def mpsa_upload(self, stiff, codes, robw, eta, alphas=()):
        self._mpsa = (stiff, codes, robw, eta, {q: a for q, a in enumerate(alphas)})
        self._basis = None

    def mpsa_set_basis(self, basis):
        self._basis = basis

    def mpsa_assemble(self):
        stiff, codes, robw, eta, al = self._mpsa
        if robw is not None:
            robw = np.asarray(robw)[:self.nd, :self.nd]
        self._mout = self.emu.mpsa(stiff, codes, robw, eta, alpha=al or None, basis=getattr(self, '_basis', None))
        return 0.0

    def mpsa_download(self):
        return {k: self._mout[k] for k in ("stress", "bound_stress", "bound_displacement_cell",
                                           "bound_displacement_face")}

    def biot_download(self, q):
        return {k: self._mout[k][q] for k in ("displacement_divergence", "boundary_displacement_divergence",
                                              "scalar_gradient", "mpsa_consistency",
                                              "bound_displacement_pressure")}


class EmuBackedFaceGrid:
    """Stand-in for ``porepy_b200.fv.FaceGrid`` backed by the host build of the per-face routines."""

    def __init__(self, sd):
        cf = sps.csc_matrix(sd.cell_faces)
        self.nc, self.nf = sd.num_cells, sd.num_faces
        self.cf = [cf.indptr.astype(np.int32), cf.indices.astype(np.int32), np.asarray(cf.data).astype(np.int8)]
        self.geo = [np.ascontiguousarray(a, np.float64) for a in (sd.face_normals, sd.face_centers, sd.cell_centers)]

    @classmethod
    def for_grid(cls, sd):
        return cls(sd)

    def _cf(self):
        return (C.c_int64(self.nc), C.c_int64(self.nf), _p(self.cf[0], C.c_int32), _p(self.cf[1], C.c_int32),
                _p(self.cf[2], C.c_int8))

    def tpfa(self, perm, bc_bits, fc_indptr, vdim):
        L = lib()
        ip = np.ascontiguousarray(fc_indptr, np.int32)
        nnz = int(ip[-1])
        nf = self.nf
        out = [np.zeros(nnz), np.zeros(nnz), np.zeros(nnz * vdim), np.zeros(nnz * vdim), np.zeros(nf), np.zeros(nf)]
        perm = np.ascontiguousarray(perm, np.float64)
        bits = np.ascontiguousarray(bc_bits, np.uint8)
        L.emu_facegrid_tpfa(*self._cf(), *[_p(a, C.c_double) for a in self.geo], _p(perm, C.c_double),
                            _p(bits, C.c_uint8), _p(ip, C.c_int32), C.c_int(vdim), *[_p(a, C.c_double) for a in out])
        return out

    def tpfa_diff(self, k_c, fc_indptr):
        L = lib()
        ip = np.ascontiguousarray(fc_indptr, np.int32)
        nhf = int(ip[-1])
        k = np.ascontiguousarray(k_c, np.float64)
        t_hf, T, dT = np.zeros(nhf), np.zeros(self.nf), np.zeros(nhf * 9)
        L.emu_facegrid_tpfa_diff(*self._cf(), *[_p(a, C.c_double) for a in self.geo], _p(k, C.c_double),
                                 _p(ip, C.c_int32), _p(t_hf, C.c_double), _p(T, C.c_double), _p(dT, C.c_double))
        return t_hf, T, dT

    def upwind(self, darcy_flux, bc_bits):
        L = lib()
        nf = self.nf
        q = np.ascontiguousarray(darcy_flux, np.float64)
        bits = np.ascontiguousarray(bc_bits, np.uint8)
        up, neu, dr = np.zeros(nf, np.int32), np.zeros(nf), np.zeros(nf)
        L.emu_facegrid_upwind(*self._cf(), _p(q, C.c_double), _p(bits, C.c_uint8), _p(up, C.c_int32),
                              _p(neu, C.c_double), _p(dr, C.c_double))
        return up, neu, dr


def face_bc_bits(bc, nf: int) -> np.ndarray:
    """Boundary byte of the per-face routines (porepy_b200/csrc/face_kernels.cuh): effective code in
    bits 0-1, raw is_dir / is_neu in bits 2 / 3."""
    internal = np.asarray(getattr(bc, "is_internal", np.zeros(nf, bool)), bool)
    is_dir, is_neu, is_rob = (np.asarray(getattr(bc, k), bool) for k in ("is_dir", "is_neu", "is_rob"))
    bits = np.zeros(nf, np.uint8)
    bits[is_rob & ~internal] = 3
    bits[is_dir & ~internal] = 1
    bits[is_neu | internal] = 2
    bits |= (is_dir.astype(np.uint8) << 2) | (is_neu.astype(np.uint8) << 3)
    return bits


def _tpfa(self, perm, bc, vdim=None):
    L = lib()
    g = self.g
    nf, nc = self.nf, self.nc
    vdim = self.nd if vdim is None else vdim
    fc = sps.csr_matrix(g.cell_faces)
    fc.sort_indices()
    ip = fc.indptr.astype(np.int32)
    nnz = fc.indices.size
    out = [np.zeros(nnz), np.zeros(nnz), np.zeros(nnz * vdim), np.zeros(nnz * vdim), np.zeros(nf), np.zeros(nf)]
    perm = np.ascontiguousarray(perm, np.float64)
    bits = face_bc_bits(bc, nf)
    L.emu_tpfa(self.h, _p(self.geo[1], C.c_double), _p(self.geo[2], C.c_double), _p(self.geo[4], C.c_double),
               _p(perm, C.c_double), _p(bits, C.c_uint8), _p(ip, C.c_int32), C.c_int(vdim),
               *[_p(a, C.c_double) for a in out])
    ix = fc.indices
    bnd = np.asarray(abs(fc).sum(axis=1)).ravel() == 1
    cols_v = (ix[:, None].astype(np.int64) * vdim + np.arange(vdim)).ravel()
    return {
        "flux": sps.csr_matrix((out[0], ix, ip), shape=(nf, nc)),
        "bound_pressure_cell": sps.csr_matrix((out[1], ix, ip), shape=(nf, nc)),
        "vector_source": sps.csr_matrix((out[2], cols_v, ip.astype(np.int64) * vdim), shape=(nf, nc * vdim)),
        "bound_pressure_vector_source": sps.csr_matrix((out[3], cols_v, ip.astype(np.int64) * vdim),
                                                       shape=(nf, nc * vdim)),
        "bound_flux": sps.diags(np.where(bnd, out[4], 0.0)).tocsr(),
        "bound_pressure_face": sps.diags(out[5]).tocsr(),
    }


def _upwind(self, darcy_flux, bc):
    L = lib()
    nf, nc = self.nf, self.nc
    q = np.ascontiguousarray(darcy_flux, np.float64)
    bits = face_bc_bits(bc, nf)
    up = np.zeros(nf, np.int32)
    neu, dr = np.zeros(nf), np.zeros(nf)
    L.emu_upwind(self.h, _p(q, C.c_double), _p(bits, C.c_uint8), _p(up, C.c_int32), _p(neu, C.c_double),
                 _p(dr, C.c_double))
    rows = np.flatnonzero(up >= 0)
    return {"upwind": sps.coo_matrix((np.ones(rows.size), (rows, up[rows])), shape=(nf, nc)).tocsr(),
            "bound_transport_neu": sps.diags(neu).tocsr(), "bound_transport_dir": sps.diags(dr).tocsr()}


EmuPlan.tpfa = _tpfa
EmuPlan.upwind = _upwind


def emu_interface_upwind_masks(interface_flux):
    """Host stand-in of ``fv.interface_upwind_masks`` (the formulas of upwind.py:496-510)."""
    s = np.sign(np.asarray(interface_flux, dtype=np.float64))
    flag = (s > 0).astype(float)
    return s, flag, 1 - flag


def geometry_3d(g):
    """Host build of the compute_geometry routines (porepy_b200/csrc/geometry_kernels.cuh) on a grid's topology and
    nodes; returns (face_normals, face_centers, face_areas, cell_centers, cell_volumes)."""
    L = lib()
    cf = sps.csc_matrix(g.cell_faces)
    fn = sps.csc_matrix(g.face_nodes)
    cf.sort_indices()
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    k = [cf.indptr.astype(np.int32), cf.indices.astype(np.int32), np.asarray(cf.data).astype(np.int8),
         fn.indptr.astype(np.int32), fn.indices.astype(np.int32)]
    nodes = np.ascontiguousarray(g.nodes, np.float64)
    out = [np.zeros((3, nf)), np.zeros((3, nf)), np.zeros(nf), np.zeros((3, nc)), np.zeros(nc)]
    rc = L.emu_geometry_3d(C.c_int64(nc), C.c_int64(nf), C.c_int64(nn), _p(k[0], C.c_int32), _p(k[1], C.c_int32),
                           _p(k[2], C.c_int8), _p(k[3], C.c_int32), _p(k[4], C.c_int32), _p(nodes, C.c_double),
                           *[_p(a, C.c_double) for a in out])
    if rc:
        raise ValueError("Some tetrahedra have negative volume")
    return out
