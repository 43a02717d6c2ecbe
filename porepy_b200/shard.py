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
    """Equal-count coordinate slabs.  Returns part id per cell."""
    cc = np.asarray(g.cell_centers)
    if axis is None:
        axis = int(np.argmax(np.ptp(cc, axis=1)))
    order = np.argsort(cc[axis], kind="stable")
    part = np.empty(g.num_cells, dtype=np.int64)
    bounds = np.linspace(0, g.num_cells, nparts + 1).astype(np.int64)
    for p in range(nparts):
        part[order[bounds[p]:bounds[p + 1]]] = p
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


def extract_cells(g, cells: np.ndarray, keep_faces: np.ndarray, keep_cells: np.ndarray, rank: int = 0) -> Shard:
    """Sub-grid of the given (sorted, global) cells as a ``Shard`` that keeps the rows of the global
    faces / cells flagged in ``keep_faces`` / ``keep_cells`` (bool per global entity).  The reference's
    ``pp.partition.extract_subgrid`` (grids/partition.py) + ``subgrid_to_grid_mapping``."""
    cf = sps.csc_matrix(g.cell_faces)
    fn = sps.csc_matrix(g.face_nodes)
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    cells = np.asarray(cells, dtype=np.int64)
    sub_cf = cf[:, cells]
    faces = np.unique(sub_cf.indices)
    sub_cf = sub_cf.tocsr()[faces].tocsc()
    sub_fn = fn[:, faces]
    nodes = np.unique(sub_fn.indices)
    sub_fn = sub_fn.tocsr()[nodes].tocsc()
    lg = Grid(g.dim, np.asarray(g.nodes)[:, nodes], sub_fn, sub_cf, name=getattr(g, "name", "Grid"))
    lg.set_geometry(np.asarray(g.face_normals)[:, faces], np.asarray(g.face_centers)[:, faces],
                    np.asarray(g.face_areas)[faces], np.asarray(g.cell_centers)[:, cells],
                    np.asarray(g.cell_volumes)[cells])
    glob_bnd = np.zeros(nf, bool)
    glob_bnd[g.get_all_boundary_faces()] = True
    loc_single = np.asarray(abs(lg.cell_faces).sum(axis=1)).ravel() == 1
    cut = loc_single & ~glob_bnd[faces]
    tags = getattr(g, "tags", {})
    if "fracture_faces" in tags:
        lg.tags["fracture_faces"] = np.asarray(tags["fracture_faces"], bool)[faces]
    lg.tags["domain_boundary_faces"] = loc_single & ~lg.tags["fracture_faces"]
    return Shard(rank, lg, cells, faces, nodes, np.asarray(keep_cells, bool)[cells],
                 np.asarray(keep_faces, bool)[faces], cut, (nc, nf, nn))


def extract_shard(g, part: np.ndarray, rank: int) -> Shard:
    """This rank's nodes, every cell touching them (one halo layer); rows of the faces of the rank's own
    cells (a shared face goes to the lower rank) and of its own cells."""
    cf = sps.csc_matrix(g.cell_faces)
    fn = sps.csc_matrix(g.face_nodes)
    nc, nf, nn = g.num_cells, g.num_faces, g.num_nodes
    own = part == rank
    cell_nodes = (abs(fn) @ abs(cf)).tocsc()  # nn x nc
    cell_nodes.data[:] = 1
    own_nodes = np.zeros(nn, bool)
    own_nodes[np.unique(cell_nodes[:, np.flatnonzero(own)].indices)] = True
    touch = np.asarray((cell_nodes.T @ own_nodes.astype(np.float64))).ravel() > 0
    # faces of own cells; shared faces go to the lower rank
    face_min_part = np.full(nf, np.iinfo(np.int64).max)
    coo = abs(cf).tocoo()
    np.minimum.at(face_min_part, coo.row, part[coo.col])
    own_face_glob = np.zeros(nf, bool)
    own_face_glob[np.unique(cf[:, np.flatnonzero(own)].indices)] = True
    own_face_glob &= face_min_part == rank
    return extract_cells(g, np.flatnonzero(touch), own_face_glob, own, rank)


def restrict_scalar_bc(bc, shard: Shard):
    """Boundary condition of the sub-grid: the global flags on true boundary faces, Neumann on
    the artificial cut faces (their rows are discarded; cf. Mpfa._bc_for_subgrid, mpfa.py:1580)."""
    from types import SimpleNamespace
    f = shard.faces
    out = SimpleNamespace(bc_type="scalar", num_faces=f.size)
    out.is_dir = np.asarray(bc.is_dir, bool)[f].copy()
    out.is_rob = np.asarray(bc.is_rob, bool)[f].copy()
    out.is_neu = np.asarray(bc.is_neu, bool)[f].copy()
    out.is_internal = np.asarray(getattr(bc, "is_internal", np.zeros(bc.is_dir.shape[-1], bool)), bool)[f].copy()
    out.robin_weight = np.asarray(bc.robin_weight, float)[f].copy()
    out.is_dir[shard.cut_face] = False
    out.is_rob[shard.cut_face] = False
    out.is_neu[shard.cut_face] = True
    return out


def restrict_vector_bc(bc, shard: Shard):
    from types import SimpleNamespace
    f = shard.faces
    out = SimpleNamespace(bc_type="vectorial", num_faces=f.size)
    out.is_dir = np.asarray(bc.is_dir, bool)[:, f].copy()
    out.is_rob = np.asarray(bc.is_rob, bool)[:, f].copy()
    out.is_neu = np.asarray(bc.is_neu, bool)[:, f].copy()
    out.is_internal = np.asarray(bc.is_internal, bool)[f].copy()
    out.robin_weight = np.asarray(bc.robin_weight, float)[:, :, f].copy()
    if getattr(bc, "basis", None) is not None:
        out.basis = np.asarray(bc.basis, float)[:, :, f].copy()
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
