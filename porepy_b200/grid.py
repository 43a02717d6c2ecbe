"""Minimal host-side mesh container and synthetic 3-D mesh generators.

``Grid`` carries exactly the arrays of ``pp.Grid`` (reference
src/porepy/grids/grid.py:32) that the discretization hot path reads, under the same
attribute names, so the discretization classes accept either a real ``pp.Grid`` or this
container (duck typing).  The generators exist because the reference (and its mesh
generators) is not present on the GPU box; their numbering is this module's own.

Not part of the hot path: everything here is O(n) NumPy run once per mesh.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sps


class Grid:
    """Topology + geometry arrays of a dim-dimensional grid embedded in 3-D.

    Attributes (shapes as in pp.Grid): ``dim``, ``nodes`` (3, nn), ``face_nodes`` csc
    (nn x nf, bool), ``cell_faces`` csc (nf x nc, +-1), ``face_normals`` / ``face_centers``
    (3, nf), ``face_areas`` (nf), ``cell_centers`` (3, nc), ``cell_volumes`` (nc),
    ``tags`` with ``domain_boundary_faces`` / ``fracture_faces`` / ``tip_faces``, ``name``.
    """

    def __init__(self, dim, nodes, face_nodes, cell_faces, name="Grid"):
        self.dim = int(dim)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.face_nodes = sps.csc_matrix(face_nodes)
        self.cell_faces = sps.csc_matrix(cell_faces)
        self.name = name
        self.num_nodes = self.nodes.shape[1]
        self.num_faces = self.face_nodes.shape[1]
        self.num_cells = self.cell_faces.shape[1]
        self.face_normals = self.face_centers = self.face_areas = None
        self.cell_centers = self.cell_volumes = None
        nf = self.num_faces
        bnd = np.asarray(abs(self.cell_faces).sum(axis=1)).ravel() == 1
        self.tags = {
            "domain_boundary_faces": bnd,
            "fracture_faces": np.zeros(nf, dtype=bool),
            "tip_faces": np.zeros(nf, dtype=bool),
        }

    # -- pp.Grid API used by the path / by callers
    def get_all_boundary_faces(self) -> np.ndarray:
        """grids/grid.py:817."""
        t = self.tags
        return np.flatnonzero(t["domain_boundary_faces"] | t["fracture_faces"] | t["tip_faces"])

    def get_boundary_faces(self) -> np.ndarray:
        return np.flatnonzero(self.tags["domain_boundary_faces"])

    def divergence(self, dim: int = 1) -> sps.csr_matrix:
        """grids/grid.py:1237: (nc*dim) x (nf*dim), rows c*dim+i, columns f*dim+i."""
        if dim == 1:
            return sps.csr_matrix(self.cell_faces.T)
        return sps.kron(self.cell_faces.T, sps.eye(dim)).tocsr()

    def compute_geometry(self):
        """pp.Grid.compute_geometry (grids/grid.py:362): 3-D grids on the device (``porepy_b200.geometry``)."""
        from .geometry import compute_geometry
        compute_geometry(self, assign=True)
        return self

    def set_geometry(self, face_normals, face_centers, face_areas, cell_centers, cell_volumes):
        self.face_normals = np.ascontiguousarray(face_normals, dtype=np.float64)
        self.face_centers = np.ascontiguousarray(face_centers, dtype=np.float64)
        self.face_areas = np.ascontiguousarray(face_areas, dtype=np.float64)
        self.cell_centers = np.ascontiguousarray(cell_centers, dtype=np.float64)
        self.cell_volumes = np.ascontiguousarray(cell_volumes, dtype=np.float64)
        return self

    @classmethod
    def from_arrays(cls, d) -> "Grid":
        """Rebuild a grid from the arrays stored by tools/make_golden.py."""
        nn = d["nodes"].shape[1]
        nf = d["fn_indptr"].size - 1
        nc = d["cf_indptr"].size - 1
        fn = sps.csc_matrix((np.ones(d["fn_indices"].size, dtype=bool), d["fn_indices"],
                             d["fn_indptr"]), shape=(nn, nf))
        cf = sps.csc_matrix((d["cf_data"].astype(np.float64), d["cf_indices"], d["cf_indptr"]),
                            shape=(nf, nc))
        g = cls(int(d["dim"]), d["nodes"], fn, cf, name=str(d["name"]))
        g.set_geometry(d["face_normals"], d["face_centers"], d["face_areas"], d["cell_centers"],
                       d["cell_volumes"])
        if "fracture_faces" in d:
            ff = np.asarray(d["fracture_faces"], bool)
            g.tags["fracture_faces"] = ff
            g.tags["domain_boundary_faces"] = g.tags["domain_boundary_faces"] & ~ff
        return g


# ----------------------------------------------------------------------------------------
# synthetic meshes (S1/S3 Cartesian, S2 structured tetrahedra)
# ----------------------------------------------------------------------------------------


def cart_grid_3d(nx, physdims=(1.0, 1.0, 1.0), perturb: float = 0.0, seed: int = 0) -> Grid:
    """Hexahedral grid of nx[0] x nx[1] x nx[2] box cells.

    ``perturb`` > 0 shifts the interior coordinate planes by perturb*h*(0.5-U): cells stay
    boxes (planar faces) but become non-uniform."""
    nx = np.asarray(nx, dtype=np.int64)
    ex, ey, ez = (int(v) for v in nx)
    xs = [np.linspace(0, physdims[i], int(nx[i]) + 1) for i in range(3)]
    if perturb > 0:
        rng = np.random.default_rng(seed)
        for i in range(3):
            h = physdims[i] / nx[i]
            xs[i][1:-1] += perturb * h * (0.5 - rng.random(int(nx[i]) - 1))
    X, Y, Z = np.meshgrid(xs[0], xs[1], xs[2], indexing="ij")
    nodes = np.vstack([a.ravel(order="F") for a in (X, Y, Z)])  # node id i + npx*(j + npy*k)
    npx, npy = ex + 1, ey + 1

    def nid(i, j, k):
        return i + npx * (j + npy * k)

    def lattice(a, b, c):
        A, B, C = np.meshgrid(np.arange(a), np.arange(b), np.arange(c), indexing="ij")
        return A.ravel(order="F"), B.ravel(order="F"), C.ravel(order="F")

    I, J, K = lattice(ex + 1, ey, ez)
    fx = np.stack([nid(I, J, K), nid(I, J + 1, K), nid(I, J + 1, K + 1), nid(I, J, K + 1)], axis=1)
    nfx = fx.shape[0]
    I2, J2, K2 = lattice(ex, ey + 1, ez)
    fy = np.stack([nid(I2, J2, K2), nid(I2, J2, K2 + 1), nid(I2 + 1, J2, K2 + 1),
                   nid(I2 + 1, J2, K2)], axis=1)
    nfy = fy.shape[0]
    I3, J3, K3 = lattice(ex, ey, ez + 1)
    fz = np.stack([nid(I3, J3, K3), nid(I3 + 1, J3, K3), nid(I3 + 1, J3 + 1, K3),
                   nid(I3, J3 + 1, K3)], axis=1)
    fnodes = np.vstack((fx, fy, fz))
    nf = fnodes.shape[0]
    nn = nodes.shape[1]
    face_nodes = sps.csc_matrix((np.ones(4 * nf, dtype=bool), fnodes.ravel(),
                                 np.arange(0, 4 * nf + 1, 4)), shape=(nn, nf))
    Ic, Jc, Kc = lattice(ex, ey, ez)
    nc = Ic.size

    def fxid(i, j, k):
        return i + (ex + 1) * (j + ey * k)

    def fyid(i, j, k):
        return nfx + i + ex * (j + (ey + 1) * k)

    def fzid(i, j, k):
        return nfx + nfy + i + ex * (j + ey * k)

    cfaces = np.stack([fxid(Ic, Jc, Kc), fxid(Ic + 1, Jc, Kc), fyid(Ic, Jc, Kc),
                       fyid(Ic, Jc + 1, Kc), fzid(Ic, Jc, Kc), fzid(Ic, Jc, Kc + 1)], axis=1)
    sgn = np.tile(np.array([-1.0, 1.0, -1.0, 1.0, -1.0, 1.0]), (nc, 1))
    cell_faces = sps.csc_matrix((sgn.ravel(), cfaces.ravel(), np.arange(0, 6 * nc + 1, 6)),
                                shape=(nf, nc))
    g = Grid(3, nodes, face_nodes, cell_faces, name="CartGrid")
    g.cart_dims = (ex, ey, ez)
    dx = [np.diff(x) for x in xs]
    xm = [0.5 * (x[1:] + x[:-1]) for x in xs]
    fc = np.zeros((3, nf))
    fnrm = np.zeros((3, nf))
    fa = np.zeros(nf)
    fc[0, :nfx], fc[1, :nfx], fc[2, :nfx] = xs[0][I], xm[1][J], xm[2][K]
    fa[:nfx] = dx[1][J] * dx[2][K]
    fnrm[0, :nfx] = fa[:nfx]
    s = slice(nfx, nfx + nfy)
    fc[0, s], fc[1, s], fc[2, s] = xm[0][I2], xs[1][J2], xm[2][K2]
    fa[s] = dx[0][I2] * dx[2][K2]
    fnrm[1, s] = fa[s]
    s = slice(nfx + nfy, nf)
    fc[0, s], fc[1, s], fc[2, s] = xm[0][I3], xm[1][J3], xs[2][K3]
    fa[s] = dx[0][I3] * dx[1][J3]
    fnrm[2, s] = fa[s]
    cc = np.vstack((xm[0][Ic], xm[1][Jc], xm[2][Kc]))
    cv = dx[0][Ic] * dx[1][Jc] * dx[2][Kc]
    return g.set_geometry(fnrm, fc, fa, cc, cv)


def cart_grid_2d(nx, physdims=(1.0, 1.0)) -> Grid:
    """Cartesian grid of nx[0] x nx[1] rectangles in the plane z = 0 (the reference's
    ``pp.CartGrid([nx, ny], physdims)``: nodes i + (nx+1) j, cells i + nx j, first the faces with
    normal along x, then those with normal along y; the stored normal has the length of the edge and
    points towards increasing x resp. y)."""
    ex, ey = int(nx[0]), int(nx[1])
    xs = np.linspace(0.0, physdims[0], ex + 1)
    ys = np.linspace(0.0, physdims[1], ey + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    nodes = np.vstack([X.ravel(order="F"), Y.ravel(order="F"), np.zeros((ex + 1) * (ey + 1))])
    npx = ex + 1

    def nid(i, j):
        return i + npx * j

    # x-faces: vertical edges at x_i between (i, j) and (i, j+1)
    I, J = np.meshgrid(np.arange(ex + 1), np.arange(ey), indexing="ij")
    I, J = I.ravel(order="F"), J.ravel(order="F")
    fx = np.stack([nid(I, J), nid(I, J + 1)], axis=1)
    nfx = fx.shape[0]
    # y-faces: horizontal edges at y_j between (i, j) and (i+1, j)
    I2, J2 = np.meshgrid(np.arange(ex), np.arange(ey + 1), indexing="ij")
    I2, J2 = I2.ravel(order="F"), J2.ravel(order="F")
    fy = np.stack([nid(I2, J2), nid(I2 + 1, J2)], axis=1)
    fnodes = np.vstack([fx, fy])
    nf = fnodes.shape[0]
    face_nodes = sps.csc_matrix((np.ones(2 * nf, dtype=bool), fnodes.ravel(), np.arange(0, 2 * nf + 1, 2)),
                                shape=(nodes.shape[1], nf))
    Ic, Jc = np.meshgrid(np.arange(ex), np.arange(ey), indexing="ij")
    Ic, Jc = Ic.ravel(order="F"), Jc.ravel(order="F")
    west = Ic + (ex + 1) * Jc
    east = west + 1
    south = nfx + Ic + ex * Jc
    north = south + ex
    cf_idx = np.stack([west, east, south, north], axis=1).ravel()
    cf_dat = np.tile(np.array([-1.0, 1.0, -1.0, 1.0]), ex * ey)
    nc = ex * ey
    cell_faces = sps.csc_matrix((cf_dat, cf_idx, np.arange(0, 4 * nc + 1, 4)), shape=(nf, nc))
    cell_faces.sort_indices()
    g = Grid(2, nodes, face_nodes, cell_faces, name="CartGrid")
    a, b = nodes[:, fnodes[:, 0]], nodes[:, fnodes[:, 1]]
    tang = b - a
    fa = np.linalg.norm(tang, axis=0)
    nrm = np.vstack([tang[1], -tang[0], np.zeros(nf)])   # rotate the edge by -90 degrees
    nrm[:, nfx:] *= -1.0                                   # y-faces: +y
    fc = 0.5 * (a + b)
    cc = np.vstack([0.5 * (xs[Ic] + xs[Ic + 1]), 0.5 * (ys[Jc] + ys[Jc + 1]), np.zeros(nc)])
    cv = (xs[Ic + 1] - xs[Ic]) * (ys[Jc + 1] - ys[Jc])
    return g.set_geometry(nrm, fc, fa, cc, cv)


def simplex_geometry_3d(g: Grid, cn: np.ndarray) -> Grid:
    """Geometry of a tetrahedral grid from its (nc,4) cell-node table.  The stored normal of a
    face points out of the cell whose ``cell_faces`` entry is +1 (pp convention)."""
    tri = g.face_nodes.indices.reshape(-1, 3)
    p = g.nodes
    a, b, c = p[:, tri[:, 0]], p[:, tri[:, 1]], p[:, tri[:, 2]]
    nrm = 0.5 * np.cross(b - a, c - a, axis=0)
    fc = (a + b + c) / 3.0
    fa = np.linalg.norm(nrm, axis=0)
    cc = (p[:, cn[:, 0]] + p[:, cn[:, 1]] + p[:, cn[:, 2]] + p[:, cn[:, 3]]) / 4.0
    e1 = p[:, cn[:, 1]] - p[:, cn[:, 0]]
    e2 = p[:, cn[:, 2]] - p[:, cn[:, 0]]
    e3 = p[:, cn[:, 3]] - p[:, cn[:, 0]]
    cv = np.abs(np.einsum("ij,ij->j", np.cross(e1, e2, axis=0), e3)) / 6.0
    cf = sps.coo_matrix(g.cell_faces)
    pos = cf.data > 0
    fpos = np.full(g.num_faces, -1, dtype=np.int64)
    fpos[cf.row[pos]] = cf.col[pos]
    fneg = np.full(g.num_faces, -1, dtype=np.int64)
    fneg[cf.row[~pos]] = cf.col[~pos]
    ref_cell = np.where(fpos >= 0, fpos, fneg)
    outward = np.einsum("ij,ij->j", nrm, fc - cc[:, ref_cell])
    flip = np.where(fpos >= 0, outward < 0, outward > 0)
    nrm[:, flip] *= -1.0
    # keep the node loop of every face consistent with its normal (right-hand rule), as pp.Grid.compute_geometry
    # derives the normal FROM the loop (grids/grid.py:590-640): swap two nodes of the flipped faces
    ix = g.face_nodes.indices.reshape(-1, 3)
    ix[flip, 1], ix[flip, 2] = ix[flip, 2].copy(), ix[flip, 1].copy()
    g.face_nodes.has_sorted_indices = False
    return g.set_geometry(nrm, fc, fa, cc, cv)


def tet_grid_from_cells(nodes: np.ndarray, cn: np.ndarray, name="TetrahedralGrid") -> Grid:
    """Faces / cell_faces of a conforming tetrahedral mesh from its (nc,4) cell-node table."""
    nc = cn.shape[0]
    nn = nodes.shape[1]
    loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    f_all = np.sort(cn[:, loc].reshape(-1, 3), axis=1)
    key = (f_all[:, 0].astype(np.int64) * nn + f_all[:, 1]) * nn + f_all[:, 2]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    nf = uniq.size
    tri = f_all[first]
    face_nodes = sps.csc_matrix((np.ones(3 * nf, dtype=bool), tri.ravel(),
                                 np.arange(0, 3 * nf + 1, 3)), shape=(nn, nf))
    order = np.argsort(inv, kind="stable")
    sgn = np.ones(4 * nc)
    so = inv[order]
    second = np.r_[False, so[1:] == so[:-1]]
    sgn[order[second]] = -1.0
    cell_faces = sps.csc_matrix((sgn, inv, np.arange(0, 4 * nc + 1, 4)), shape=(nf, nc))
    cell_faces.sort_indices()
    g = Grid(3, nodes, face_nodes, cell_faces, name=name)
    return simplex_geometry_3d(g, cn)


def structured_tet_grid(nx, physdims=(1.0, 1.0, 1.0)) -> Grid:
    """Each box of an nx[0] x nx[1] x nx[2] lattice is split into 6 tetrahedra around its
    main diagonal (Kuhn triangulation; conforming across boxes): 6*prod(nx) cells."""
    nx = np.asarray(nx, dtype=np.int64)
    ex, ey, ez = (int(v) for v in nx)
    xs = [np.linspace(0, physdims[i], int(nx[i]) + 1) for i in range(3)]
    X, Y, Z = np.meshgrid(xs[0], xs[1], xs[2], indexing="ij")
    nodes = np.vstack([a.ravel(order="F") for a in (X, Y, Z)])
    npx, npy = ex + 1, ey + 1
    Ic, Jc, Kc = np.meshgrid(np.arange(ex), np.arange(ey), np.arange(ez), indexing="ij")
    Ic, Jc, Kc = (a.ravel(order="F") for a in (Ic, Jc, Kc))

    def nid(d):
        return (Ic + d[0]) + npx * ((Jc + d[1]) + npy * (Kc + d[2]))

    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=int)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] = 1
            v.append(w)
        tets.append(np.stack([nid(vv) for vv in v], axis=1))
    cn = np.stack(tets, axis=1).reshape(-1, 4)  # the 6 tets of a box are consecutive cells
    return tet_grid_from_cells(nodes, cn, name="StructuredTetrahedralGrid")
