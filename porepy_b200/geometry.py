"""``Grid.compute_geometry`` on the device for 3-D grids (reference src/porepy/grids/grid.py:362-381 dispatch,
:572-778 ``_compute_geometry_3d``): face normals / centres / areas and cell centres / volumes from the topology and
the nodes, through ``pb_compute_geometry_3d`` (csrc/geometry.cu: one thread per face, then one per cell).

Lower-dimensional grids (fracture planes, intersection lines, points: < 1 % of the cells of a mixed-dimensional
grid) keep the reference's own ``compute_geometry``; asking for them here raises ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sps

from . import _lib


def compute_geometry(g, assign: bool = True):
    """Compute (face_normals, face_centers, face_areas, cell_centers, cell_volumes) of the 3-D grid ``g`` (a
    ``pp.Grid`` or ``porepy_b200.Grid``: ``dim``, ``nodes``, ``face_nodes`` with the nodes of every face in loop order,
    ``cell_faces``).  ``assign``: also store them on ``g`` under the reference's attribute names.  Raises
    ``ValueError("Some tetrahedra have negative volume")`` as the reference does (grid.py:754)."""
    if int(g.dim) != 3:
        raise NotImplementedError("porepy_b200.compute_geometry: 3-D grids only (the reference handles dim < 3)")
    _lib.require_gpu()
    lib = _lib.load()
    cf = sps.csc_matrix(g.cell_faces)
    if not cf.has_sorted_indices:
        cf = cf.copy()
        cf.sort_indices()           # the reference sums the sub-tetrahedra of a cell in ascending face order
    fn = g.face_nodes if sps.isspmatrix_csc(g.face_nodes) else sps.csc_matrix(g.face_nodes)
    nc, nf, nn = cf.shape[1], cf.shape[0], fn.shape[0]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
    cf_ip, cf_ix, cf_da = i32(cf.indptr), i32(cf.indices), np.ascontiguousarray(cf.data, dtype=np.int8)
    fn_ip, fn_ix = i32(fn.indptr), i32(fn.indices)
    nodes = np.ascontiguousarray(np.asarray(g.nodes, dtype=np.float64)[:3])
    out = [np.empty((3, nf)), np.empty((3, nf)), np.empty(nf), np.empty((3, nc)), np.empty(nc)]
    ms = C.c_float()
    _lib.check(lib.pb_compute_geometry_3d(nc, nf, nn, _lib.ptr(cf_ip, _lib._i32p), _lib.ptr(cf_ix, _lib._i32p),
                                          _lib.ptr(cf_da, _lib._i8p), _lib.ptr(fn_ip, _lib._i32p),
                                          _lib.ptr(fn_ix, _lib._i32p), _lib.ptr(nodes, _lib._f64p),
                                          *[_lib.ptr(a, _lib._f64p) for a in out], C.byref(ms)))
    compute_geometry.last_kernel_ms = float(ms.value)
    if assign:
        g.face_normals, g.face_centers, g.face_areas, g.cell_centers, g.cell_volumes = out
    return tuple(out)


compute_geometry.last_kernel_ms = 0.0
