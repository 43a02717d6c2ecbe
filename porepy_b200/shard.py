"""Sharding of one grid's interaction regions across GPUs (one process per GPU).

The assembly needs no data-path collective: interaction regions (nodes) are independent and a
face row is complete as soon as all nodes of the face have been processed.  The decomposition
is the reference's own memory-splitting scheme (``_fvutils.subproblems``, reference
src/porepy/numerics/fv/_fvutils.py:414-539; overlap rows removed by
``remove_nonlocal_contribution`` :542, mpfa.py:301-304) with the simplification that every face
row is produced by exactly one shard (no averaging over repeated faces, mpfa.py:357-372):

* cells are partitioned (coordinate slabs, the structured analogue of
  ``pp.partition.partition``, grids/partition.py:269);
* shard p takes the NODES of its cells and every cell touching those nodes (one halo layer):
  all interaction regions of the faces of its own cells are then complete;
* shard p keeps the rows of the faces of its own cells; a face shared by two shards is kept
  by the lower rank.  Cell rows (Biot's divergence-type terms) are kept for own cells.

``Shard.to_global`` maps the kept rows into global numbering; concatenating (summing) the
shards' matrices reproduces the unsplit discretization (tests/test_shard*.py; parity with
applications/test_utils/common_xpfa_tests.py:832-957 "split == unsplit").
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sps

from .grid import Grid


def partition_cells(g, nparts: int, axis: int | None = None) -> np.ndarray:
    """Part id per cell.  ``axis`` given: equal-count coordinate slabs along it.  Default: recursive coordinate
    bisection (equal counts, always across the longest extent of the current piece) -- compact blocks, so the
    one-layer halo of ``extract_shard`` stays small (2 x 2 x 2 blocks of a cube at 8 parts: ~11 % extra regions
    at 10^6 tetrahedra, against ~29 % for 8 slabs).  The structured analogue of ``pp.partition.partition``
    (grids/partition.py:269)."""
    cc = np.asarray(g.cell_centers)
    part = np.empty(g.num_cells, dtype=np.int64)
    if axis is not None:
        order = np.argsort(cc[axis], kind="stable")
        bounds = np.linspace(0, g.num_cells, nparts + 1).astype(np.int64)
        for p in range(nparts):
            part[order[bounds[p]:bounds[p + 1]]] = p
        return part

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        x = cc[:, idx]
        ax = int(np.argmax(np.ptp(x, axis=1)))
        left = count // 2
        cut = (idx.size * left) // count
        order = np.argsort(x[ax], kind="stable")
        split(idx[order[:cut]], first, left)
        split(idx[order[cut:]], first + left, count - left)

    split(np.arange(g.num_cells), 0, int(nparts))
    return part


@dataclass
class Shard:
    rank: int
    grid: Grid              # local sub-grid (own cells + one halo layer)
    cells: np.ndarray       # local -> global cell
    faces: np.ndarray       # local -> global face
    nodes: np.ndarray       # local -> global node
    own_cell: np.ndarray    # bool per local cell
    own_face: np.ndarray    # bool per local face: rows this shard keeps
    cut_face: np.ndarray    # bool per local face: artificial boundary of the overlap
    num_global: tuple       # (nc, nf, nn) of the global grid
    own_node: np.ndarray | None = None   # bool per local node: regions this shard must assemble (None: all)

    def restrict_cell_array(self, a: np.ndarray) -> np.ndarray:
        """(..., nc_global) -> (..., nc_local)."""
        return np.ascontiguousarray(a[..., self.cells])

    def restrict_face_array(self, a: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(a[..., self.faces])

    def to_global(self, m, rows: str, cols: str, br: int = 1, bc: int = 1) -> sps.csr_matrix:
        """Embed a local matrix into global numbering, keeping only this shard's rows.
        rows/cols in {"face", "cell"}; br/bc = block sizes (nd for vector quantities)."""
        nc, nf, _ = self.num_global
        rmap, rkeep, nr = (self.faces, self.own_face, nf) if rows == "face" else (self.cells, self.own_cell, nc)
        cmap, ncg = (self.faces, nf) if cols == "face" else (self.cells, nc)
        m = sps.coo_matrix(m)
        lr, li = np.divmod(m.row, br)
        lc, lj = np.divmod(m.col, bc)
        keep = rkeep[lr]
        return sps.coo_matrix((m.data[keep], (rmap[lr[keep]] * br + li[keep], cmap[lc[keep]] * bc + lj[keep])),
                              shape=(nr * br, ncg * bc)).tocsr()


def _ranges(indptr: np.ndarray, idx: np.ndarray):
    """Positions of the concatenated slices ``indptr[i]:indptr[i+1]``, i in ``idx``, and their lengths."""
    starts = indptr[idx].astype(np.int64)
    lens = indptr[idx + 1].astype(np.int64) - starts
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64), lens
    ends = np.cumsum(lens)
    pos = np.arange(total, dtype=np.int64) + np.repeat(starts - (ends - lens), lens)
    return pos, lens


def _sub_csc(m: sps.csc_matrix, cols: np.ndarray, row_map: np.ndarray, nrows: int):
    """Columns ``cols`` of ``m`` with rows renumbered by ``row_map`` (every row of the selected columns must
    be mapped): index arithmetic only, no scipy slicing."""
    pos, lens = _ranges(m.indptr, cols)
    indptr = np.zeros(cols.size + 1, dtype=np.int32)
    np.cumsum(lens, out=indptr[1:])
    return sps.csc_matrix((m.data[pos], row_map[m.indices[pos]].astype(np.int32), indptr), shape=(nrows, cols.size))


def extract_cells(g, cells: np.ndarray, keep_faces: np.ndarray, keep_cells: np.ndarray, rank: int = 0) -> Shard:
    """Sub-grid of the given (sorted, global) cells as a ``Shard`` that keeps the rows of the global
    faces / cells flagged in ``keep_faces`` / ``keep_cells`` (bool per global entity).  The reference's
    ``pp.partition.extract_subgrid`` (grids/partition.py) + ``subgrid_to_grid_mapping``; here by masks and
    index arithmetic on the CSC arrays (O(size of the global arrays), no sparse-matrix slicing)."""
    cf = sps.csc_matrix(g.cell_faces)
    fn = sps.csc_matrix(g.face_nodes)
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    cells = np.asarray(cells, dtype=np.int64)
    pos, _ = _ranges(cf.indptr, cells)
    fmask = np.zeros(nf, bool)
    fmask[cf.indices[pos]] = True
    faces = np.flatnonzero(fmask)
    posn, _ = _ranges(fn.indptr, faces)
    nmask = np.zeros(nn, bool)
    nmask[fn.indices[posn]] = True
    nodes = np.flatnonzero(nmask)
    fmap = np.cumsum(fmask) - 1
    nmap = np.cumsum(nmask) - 1
    sub_cf = _sub_csc(cf, cells, fmap, faces.size)
    sub_fn = _sub_csc(fn, faces, nmap, nodes.size)
    lg = Grid(g.dim, np.asarray(g.nodes)[:, nodes], sub_fn, sub_cf, name=getattr(g, "name", "Grid"))
    lg.set_geometry(np.asarray(g.face_normals)[:, faces], np.asarray(g.face_centers)[:, faces],
                    np.asarray(g.face_areas)[faces], np.asarray(g.cell_centers)[:, cells],
                    np.asarray(g.cell_volumes)[cells])
    glob_count = np.bincount(cf.indices, minlength=nf)
    loc_count = np.bincount(sub_cf.indices, minlength=faces.size)
    loc_single = loc_count == 1
    cut = loc_single & (glob_count[faces] != 1)
    tags = getattr(g, "tags", {})
    if "fracture_faces" in tags:
        lg.tags["fracture_faces"] = np.asarray(tags["fracture_faces"], bool)[faces]
    lg.tags["domain_boundary_faces"] = loc_single & ~lg.tags["fracture_faces"]
    return Shard(rank, lg, cells, faces, nodes, np.asarray(keep_cells, bool)[cells],
                 np.asarray(keep_faces, bool)[faces], cut, (nc, nf, nn))


def shard_cells(g, part: np.ndarray, rank: int):
    """(cells of the shard, kept-face flags, own-cell flags) of ``rank``: its own cells' NODES, every cell touching
    one of them (one halo layer); rows of the faces of the rank's own cells (a shared face goes to the lower rank)
    and of its own cells.  Masks and segmented reductions over the CSC arrays only."""
    cf = sps.csc_matrix(g.cell_faces)
    fn = sps.csc_matrix(g.face_nodes)
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    own = np.asarray(part) == rank
    own_cells = np.flatnonzero(own)
    pos, _ = _ranges(cf.indptr, own_cells)
    own_face = np.zeros(nf, bool)
    own_face[cf.indices[pos]] = True
    posn, _ = _ranges(fn.indptr, np.flatnonzero(own_face))
    own_node = np.zeros(nn, bool)
    own_node[fn.indices[posn]] = True
    # a cell touches a node iff one of its faces contains it (every vertex of a cell lies on nd of its faces)
    face_touch = np.logical_or.reduceat(own_node[fn.indices], fn.indptr[:-1].astype(np.int64))
    face_touch[np.diff(fn.indptr) == 0] = False
    touch = np.logical_or.reduceat(face_touch[cf.indices], cf.indptr[:-1].astype(np.int64))
    touch[np.diff(cf.indptr) == 0] = False
    # shared faces go to the lower rank
    face_min_part = np.full(nf, np.iinfo(np.int64).max)
    cell_of_entry = np.repeat(np.arange(nc, dtype=np.int64), np.diff(cf.indptr))
    np.minimum.at(face_min_part, cf.indices, np.asarray(part, dtype=np.int64)[cell_of_entry])
    own_face &= face_min_part == rank
    return np.flatnonzero(touch), own_face, own


def extract_shard(g, part: np.ndarray, rank: int) -> Shard:
    """This rank's nodes, every cell touching them (one halo layer); rows of the faces of the rank's own
    cells (a shared face goes to the lower rank) and of its own cells.  Local cell numbering: OWN CELLS FIRST
    (ascending global id), then the halo cells -- the rows of a system matrix assembled on the shard that
    belong to this rank are then a prefix, and its columns are already [own | ghost] (``krylov``).

    Runs in the native library (``pb_shard_create``, csrc/shard.cu: one pass over the global CSC arrays, ~20x the
    speed of the NumPy restatement ``extract_shard_numpy`` below, which the tests hold it against)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    cf = g.cell_faces if sps.isspmatrix_csc(g.cell_faces) else sps.csc_matrix(g.cell_faces)
    fn = g.face_nodes if sps.isspmatrix_csc(g.face_nodes) else sps.csc_matrix(g.face_nodes)
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
    cf_ip, cf_ix, cf_da = i32(cf.indptr), i32(cf.indices), np.ascontiguousarray(cf.data, dtype=np.float64)
    fn_ip, fn_ix = i32(fn.indptr), i32(fn.indices)
    part64 = np.ascontiguousarray(part, dtype=np.int64)
    h = C.c_void_p()
    _lib.check(lib.pb_shard_create(nc, nf, nn, _lib.ptr(cf_ip, _lib._i32p), _lib.ptr(cf_ix, _lib._i32p),
                                   _lib.ptr(cf_da, _lib._f64p), _lib.ptr(fn_ip, _lib._i32p), _lib.ptr(fn_ix, _lib._i32p),
                                   _lib.ptr(part64, _lib._i64p), int(rank), C.byref(h)))
    try:
        sz = np.zeros(6, dtype=np.int64)
        _lib.check(lib.pb_shard_sizes(h, _lib.ptr(sz, _lib._i64p)))
        ncl, nfl, nnl, n_own, nnz_cf, nnz_fn = (int(v) for v in sz)
        cells, faces, nodes = (np.empty(n, dtype=np.int64) for n in (ncl, nfl, nnl))
        own_face, cut, single = (np.empty(nfl, dtype=np.uint8) for _ in range(3))
        own_node = np.empty(nnl, dtype=np.uint8)
        l_cf_ip, l_cf_ix, l_cf_da = np.empty(ncl + 1, np.int32), np.empty(nnz_cf, np.int32), np.empty(nnz_cf, np.float64)
        l_fn_ip, l_fn_ix = np.empty(nfl + 1, np.int32), np.empty(nnz_fn, np.int32)
        _lib.check(lib.pb_shard_fill(h, _lib.ptr(cells, _lib._i64p), _lib.ptr(faces, _lib._i64p), _lib.ptr(nodes, _lib._i64p),
                                     _lib.ptr(own_face, _lib._u8p), _lib.ptr(cut, _lib._u8p), _lib.ptr(single, _lib._u8p),
                                     _lib.ptr(own_node, _lib._u8p), _lib.ptr(l_cf_ip, _lib._i32p),
                                     _lib.ptr(l_cf_ix, _lib._i32p), _lib.ptr(l_cf_da, _lib._f64p),
                                     _lib.ptr(l_fn_ip, _lib._i32p), _lib.ptr(l_fn_ix, _lib._i32p)))
    finally:
        lib.pb_shard_destroy(h)

    def gather(a, idx):
        a = np.ascontiguousarray(a, dtype=np.float64)
        a2 = a.reshape(-1, a.shape[-1])
        out = np.empty((a2.shape[0], idx.size), dtype=np.float64)
        _lib.check(lib.pb_gather_columns(_lib.ptr(a2, _lib._f64p), a2.shape[0], a2.shape[1], _lib.ptr(idx, _lib._i64p),
                                         idx.size, _lib.ptr(out, _lib._f64p)))
        return out.reshape(a.shape[:-1] + (idx.size,))

    sub_cf = sps.csc_matrix((l_cf_da, l_cf_ix, l_cf_ip), shape=(nfl, ncl))
    sub_fn = sps.csc_matrix((np.ones(nnz_fn, dtype=bool), l_fn_ix, l_fn_ip), shape=(nnl, nfl))
    lg = Grid.__new__(Grid)
    lg.dim, lg.name = int(g.dim), getattr(g, "name", "Grid")
    lg.nodes, lg.face_nodes, lg.cell_faces = gather(g.nodes, nodes), sub_fn, sub_cf
    lg.num_nodes, lg.num_faces, lg.num_cells = nnl, nfl, ncl
    lg.set_geometry(gather(g.face_normals, faces), gather(g.face_centers, faces), gather(g.face_areas, faces),
                    gather(g.cell_centers, cells), gather(g.cell_volumes, cells))
    single = single.astype(bool)
    tags = getattr(g, "tags", {})
    frac = np.asarray(tags["fracture_faces"], bool)[faces] if "fracture_faces" in tags else np.zeros(nfl, bool)
    lg.tags = {"domain_boundary_faces": single & ~frac, "fracture_faces": frac, "tip_faces": np.zeros(nfl, bool)}
    own_cell = np.zeros(ncl, dtype=bool)
    own_cell[:n_own] = True
    return Shard(rank, lg, cells, faces, nodes, own_cell, own_face.astype(bool), cut.astype(bool), (nc, nf, nn),
                 own_node.astype(bool))


def extract_shard_numpy(g, part: np.ndarray, rank: int) -> Shard:
    """NumPy restatement of ``extract_shard`` (masks and segmented reductions over the CSC arrays): the checker of
    the native routine in the tests."""
    cells, own_face, own = shard_cells(g, part, rank)
    cells = np.concatenate((cells[own[cells]], cells[~own[cells]]))
    s = extract_cells(g, cells, own_face, own, rank)
    # nodes of the own cells: the interaction regions this shard needs (faces of own cells -> their nodes)
    lcf = sps.csc_matrix(s.grid.cell_faces)
    lfn = sps.csc_matrix(s.grid.face_nodes)
    n_own = int(s.own_cell.sum())
    fmask = np.zeros(s.grid.num_faces, bool)
    fmask[lcf.indices[:lcf.indptr[n_own]]] = True
    pos, _ = _ranges(lfn.indptr, np.flatnonzero(fmask))
    s.own_node = np.zeros(s.grid.num_nodes, bool)
    s.own_node[lfn.indices[pos]] = True
    return s


def _take_faces(a, idx: np.ndarray, dtype) -> np.ndarray:
    """``a[..., idx]`` as a fresh contiguous array: float64 arrays through the native threaded gather
    (``pb_gather_columns``; NumPy's fancy indexing of a (3, 3, nf) array along its last axis costs ~0.1 s at 2 * 10^6
    faces), flags through ``np.take``."""
    a = np.asarray(a)
    if dtype is float and a.ndim >= 1 and a.size:
        from . import _lib
        lib = _lib.load()
        a = np.ascontiguousarray(a, dtype=np.float64)
        a2 = a.reshape(-1, a.shape[-1])
        ix = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty((a2.shape[0], ix.size), dtype=np.float64)
        _lib.check(lib.pb_gather_columns(_lib.ptr(a2, _lib._f64p), a2.shape[0], a2.shape[1], _lib.ptr(ix, _lib._i64p),
                                         ix.size, _lib.ptr(out, _lib._f64p)))
        return out.reshape(a.shape[:-1] + (ix.size,))
    return np.take(np.asarray(a, dtype=dtype), idx, axis=-1)


def restrict_scalar_bc(bc, shard: Shard):
    """Boundary condition of the sub-grid: the global flags on true boundary faces, Neumann on
    the artificial cut faces (their rows are discarded; cf. Mpfa._bc_for_subgrid, mpfa.py:1580)."""
    from types import SimpleNamespace
    f = shard.faces
    out = SimpleNamespace(bc_type="scalar", num_faces=f.size)
    out.is_dir = _take_faces(bc.is_dir, f, bool)
    out.is_rob = _take_faces(bc.is_rob, f, bool)
    out.is_neu = _take_faces(bc.is_neu, f, bool)
    out.is_internal = _take_faces(getattr(bc, "is_internal", np.zeros(np.asarray(bc.is_dir).shape[-1], bool)), f, bool)
    out.robin_weight = _take_faces(bc.robin_weight, f, float)
    out.is_dir[shard.cut_face] = False
    out.is_rob[shard.cut_face] = False
    out.is_neu[shard.cut_face] = True
    return out


def restrict_vector_bc(bc, shard: Shard):
    from types import SimpleNamespace
    f = shard.faces
    out = SimpleNamespace(bc_type="vectorial", num_faces=f.size)
    out.is_dir = _take_faces(bc.is_dir, f, bool)
    out.is_rob = _take_faces(bc.is_rob, f, bool)
    out.is_neu = _take_faces(bc.is_neu, f, bool)
    out.is_internal = _take_faces(bc.is_internal, f, bool)
    # Robin weights and rotated bases act on boundary faces only: gather the (3, 3, nf) arrays only when the shard has a
    # Robin face / a boundary face whose basis is not the identity (a 10^6-cell shard otherwise moves 2 x 75 MB for nothing)
    rw = np.asarray(bc.robin_weight, float)
    if out.is_rob.any():
        out.robin_weight = _take_faces(rw, f, float)
    else:
        out.robin_weight = np.broadcast_to(np.zeros(rw.shape[:-1] + (1,)), rw.shape[:-1] + (f.size,))
    basis = getattr(bc, "basis", None)
    if basis is not None:
        basis = np.asarray(basis, float)
        bnd = f[np.asarray(shard.grid.tags["domain_boundary_faces"], bool) | np.asarray(shard.grid.tags["fracture_faces"], bool)]
        nd = basis.shape[0]
        if np.array_equal(basis[:, :, bnd], np.broadcast_to(np.eye(nd)[:, :, None], (nd, nd, bnd.size))):
            out.basis = np.broadcast_to(np.eye(nd)[:, :, None], (nd, nd, f.size))
        else:
            out.basis = _take_faces(basis, f, float)
    out.is_dir[:, shard.cut_face] = False
    out.is_rob[:, shard.cut_face] = False
    out.is_neu[:, shard.cut_face] = True
    return out


# ------------------------------------------------------------------------------------------
# one call per rank: restrict the parameters, discretize the shard, embed the kept rows
# ------------------------------------------------------------------------------------------

# key -> (row entity, column entity); block sizes follow from the local matrix shape
_LAYOUT = {
    "flux": ("face", "cell"), "bound_flux": ("face", "face"),
    "bound_pressure_cell": ("face", "cell"), "bound_pressure_face": ("face", "face"),
    "vector_source": ("face", "cell"), "bound_pressure_vector_source": ("face", "cell"),
    "stress": ("face", "cell"), "bound_stress": ("face", "face"),
    "bound_displacement_cell": ("face", "cell"), "bound_displacement_face": ("face", "face"),
    "displacement_divergence": ("cell", "cell"), "boundary_displacement_divergence": ("cell", "face"),
    "scalar_gradient": ("face", "cell"), "mpsa_consistency": ("cell", "cell"),
    "bound_displacement_pressure": ("face", "cell"),
}


def embed(shard: Shard, key: str, m):
    """Kept rows of the local matrix ``m`` (or dict of matrices) of discretization term ``key`` in
    global numbering."""
    if isinstance(m, dict):
        return {k: embed(shard, key, v) for k, v in m.items()}
    rows, cols = _LAYOUT[key]
    n_r = shard.grid.num_faces if rows == "face" else shard.grid.num_cells
    n_c = shard.grid.num_faces if cols == "face" else shard.grid.num_cells
    return shard.to_global(m, rows, cols, m.shape[0] // n_r, m.shape[1] // n_c)


def restrict_parameters(params: dict, shard: Shard) -> dict:
    """Parameter dictionary of the shard's sub-grid: cell tensors restricted to its cells, boundary
    conditions to its faces (cut faces -> Neumann), everything else passed through."""
    from .params import FourthOrderTensor, SecondOrderTensor
    out = {}
    for key, val in params.items():
        if key in ("specified_cells", "specified_faces", "specified_nodes", "active_cells", "active_faces",
                   "update_discretization"):
            continue  # describe the GLOBAL grid; the sub-grid is discretized as a whole
        if key == "second_order_tensor":
            out[key] = SecondOrderTensor.from_values(shard.restrict_cell_array(val.values))
        elif key == "fourth_order_tensor":
            out[key] = FourthOrderTensor.from_values(shard.restrict_cell_array(val.values))
        elif key == "bc":
            vec = np.asarray(val.is_dir).ndim == 2
            out[key] = restrict_vector_bc(val, shard) if vec else restrict_scalar_bc(val, shard)
        elif key == "scalar_vector_mappings":
            out[key] = {k: (SecondOrderTensor.from_values(shard.restrict_cell_array(a.values))
                            if hasattr(a, "values") else a) for k, a in val.items()}
        elif key == "bc_values":
            arr = np.asarray(val)
            nd = arr.size // shard.num_global[1]
            out[key] = arr.reshape(-1, nd)[shard.faces].ravel() if nd > 1 else arr[shard.faces]
        else:
            out[key] = val
    return out


def discretize_shard(discr, g, data: dict, part: np.ndarray, rank: int) -> dict:
    """Discretize this rank's share of ``g`` with ``discr`` (``Mpfa``, ``Mpsa`` or ``Biot``): the
    interaction regions of the rank's nodes on its sub-grid, then the rows of its own faces /
    cells in GLOBAL numbering.  Returns ``{key: csr}`` (Biot's coupling keys: ``{key: {kw: csr}}``)
    of global shape holding only this rank's rows; the sum over ranks is the unsplit discretization.
    No communication.  ``data`` is the global data dictionary (``initialize_data``); the global
    ``mpfa_eta`` / ``mpsa_eta`` is fixed from the GLOBAL grid so that all shards agree."""
    from .fv import determine_eta
    from .params import DISCRETIZATION_MATRICES, PARAMETERS, initialize_data
    kw = discr.keyword
    shard = extract_shard(g, part, rank)
    params = restrict_parameters(data[PARAMETERS][kw], shard)
    eta_key = "mpfa_eta" if "second_order_tensor" in params else "mpsa_eta"
    params.setdefault(eta_key, determine_eta(g))
    local = initialize_data({}, kw, params)
    if shard.own_node is not None and shard.grid.dim >= 2:
        from .fv import DevicePlan
        plan = DevicePlan.for_grid(shard.grid)
        if hasattr(plan, "set_active_nodes"):
            plan.set_active_nodes(shard.own_node)
    discr.discretize(shard.grid, local)
    return {key: embed(shard, key, m) for key, m in local[DISCRETIZATION_MATRICES][kw].items()}


def sum_shards(parts: list) -> dict:
    """Combine the per-rank results of ``discretize_shard`` (e.g. after ``gather_object``)."""
    acc: dict = {}
    for p in parts:
        for key, m in p.items():
            if isinstance(m, dict):
                d = acc.setdefault(key, {})
                for k, v in m.items():
                    d[k] = v if k not in d else d[k] + v
            else:
                acc[key] = m if key not in acc else acc[key] + m
    return acc
