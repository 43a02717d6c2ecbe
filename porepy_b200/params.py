"""Parameter containers mirroring the reference's (read-only inputs of the hot path).

Same attribute names and array layouts as ``pp.SecondOrderTensor`` / ``pp.FourthOrderTensor``
(src/porepy/params/tensor.py:68,251), ``pp.BoundaryCondition`` / ``BoundaryConditionVectorial``
(src/porepy/params/bc.py:68,222) and ``pp.initialize_data`` (src/porepy/params/data.py:116),
so the discretization classes take either these or the reference's objects (duck typing).
They exist because the reference is not importable on the GPU box.
"""
from __future__ import annotations

import numpy as np

PARAMETERS = "parameters"
DISCRETIZATION_MATRICES = "discretization_matrices"


class SecondOrderTensor:
    """``values`` is (3,3,nc) also in 2-D (unit kyy/kzz defaults, tensor.py:109-112)."""

    def __init__(self, kxx, kyy=None, kzz=None, kxy=None, kxz=None, kyz=None):
        kxx = np.atleast_1d(np.asarray(kxx, dtype=float))
        nc = kxx.size
        z = np.zeros(nc)
        kyy = kxx if kyy is None else np.asarray(kyy, float)
        kzz = kxx if kzz is None else np.asarray(kzz, float)
        kxy = z if kxy is None else np.asarray(kxy, float)
        kxz = z if kxz is None else np.asarray(kxz, float)
        kyz = z if kyz is None else np.asarray(kyz, float)
        v = np.zeros((3, 3, nc))
        v[0, 0], v[1, 1], v[2, 2] = kxx, kyy, kzz
        v[0, 1] = v[1, 0] = kxy
        v[0, 2] = v[2, 0] = kxz
        v[1, 2] = v[2, 1] = kyz
        self.values = v

    @classmethod
    def from_values(cls, values):
        t = cls.__new__(cls)
        t.values = np.ascontiguousarray(values, dtype=float)
        return t


class FourthOrderTensor:
    """Isotropic stiffness from Lame parameters; ``values`` is (9,9,nc) with row index
    p = 3*i + r <-> sigma_ir and column q = 3*a + k <-> du_a/dx_k (tensor.py:303-348)."""

    def __init__(self, mu, lmbda):
        mu = np.atleast_1d(np.asarray(mu, dtype=float))
        lmbda = np.atleast_1d(np.asarray(lmbda, dtype=float))
        nc = mu.size
        self.mu, self.lmbda = mu, lmbda
        v = np.zeros((9, 9, nc))
        for i in range(3):
            for r in range(3):
                p = 3 * i + r
                # sigma_ir = mu (du_i/dx_r + du_r/dx_i) + lambda delta_ir div u
                v[p, 3 * i + r] += mu
                v[p, 3 * r + i] += mu
                if i == r:
                    for a in range(3):
                        v[p, 3 * a + a] += lmbda
        self.values = v

    @classmethod
    def from_values(cls, values):
        t = cls.__new__(cls)
        t.values = np.ascontiguousarray(values, dtype=float)
        return t


class BoundaryCondition:
    """Scalar BC: face flags ``is_dir`` / ``is_neu`` / ``is_rob`` / ``is_internal``,
    ``robin_weight`` (nf).  Boundary faces default to Neumann (bc.py:130-140)."""

    bc_type = "scalar"

    def __init__(self, sd, faces=None, cond=None):
        nf = sd.num_faces
        self.num_faces = nf
        self.dim = sd.dim - 1
        self.is_neu = np.zeros(nf, dtype=bool)
        self.is_dir = np.zeros(nf, dtype=bool)
        self.is_rob = np.zeros(nf, dtype=bool)
        self.is_internal = np.asarray(sd.tags["fracture_faces"], dtype=bool).copy()
        self.is_neu[sd.get_all_boundary_faces()] = True
        self.robin_weight = np.ones(nf)
        self.basis = np.ones(nf)
        if faces is not None:
            faces = np.asarray(faces)
            if faces.dtype == bool:
                faces = np.flatnonzero(faces)
            if isinstance(cond, str):
                cond = [cond] * faces.size
            for f, c in zip(faces, cond):
                c = c.lower()
                self.is_neu[f] = self.is_dir[f] = self.is_rob[f] = False
                if c in ("dir", "dirichlet"):
                    self.is_dir[f] = True
                elif c in ("neu", "neumann"):
                    self.is_neu[f] = True
                elif c in ("rob", "robin"):
                    self.is_rob[f] = True
                else:
                    raise ValueError(f"Boundary should be Dirichlet, Neumann or Robin, not {c}")


class BoundaryConditionVectorial:
    """Vector BC: flags (nd, nf) per component; ``robin_weight`` / ``basis`` (nd,nd,nf)
    (bc.py:222-322)."""

    bc_type = "vectorial"

    def __init__(self, sd, faces=None, cond=None):
        nf, nd = sd.num_faces, sd.dim
        self.num_faces = nf
        self.dim = nd
        self.is_neu = np.zeros((nd, nf), dtype=bool)
        self.is_dir = np.zeros((nd, nf), dtype=bool)
        self.is_rob = np.zeros((nd, nf), dtype=bool)
        self.is_internal = np.asarray(sd.tags["fracture_faces"], dtype=bool).copy()
        self.is_neu[:, sd.get_all_boundary_faces()] = True
        self.robin_weight = np.tile(np.eye(nd)[:, :, None], (1, 1, nf))
        self.basis = np.tile(np.eye(nd)[:, :, None], (1, 1, nf))
        if faces is not None:
            faces = np.asarray(faces)
            if faces.dtype == bool:
                faces = np.flatnonzero(faces)
            if isinstance(cond, str):
                cond = [cond] * faces.size
            for f, c in zip(faces, cond):
                c = c.lower()
                self.is_neu[:, f] = self.is_dir[:, f] = self.is_rob[:, f] = False
                if c in ("dir", "dirichlet"):
                    self.is_dir[:, f] = True
                elif c in ("neu", "neumann"):
                    self.is_neu[:, f] = True
                elif c in ("rob", "robin"):
                    self.is_rob[:, f] = True
                else:
                    raise ValueError(f"Boundary should be Dirichlet, Neumann or Robin, not {c}")

    def internal_to_dirichlet(self, sd) -> None:
        """bc.py:53-66."""
        ff = np.asarray(sd.tags["fracture_faces"], bool)
        self.is_neu[:, ff] = False
        self.is_dir[:, ff] = True


def initialize_data(data: dict, keyword: str, specified_parameters: dict | None = None) -> dict:
    """params/data.py:116 -- put parameters under data["parameters"][keyword] and make
    sure data["discretization_matrices"][keyword] exists."""
    data.setdefault(PARAMETERS, {})
    data[PARAMETERS].setdefault(keyword, {})
    data[PARAMETERS][keyword].update(specified_parameters or {})
    data.setdefault(DISCRETIZATION_MATRICES, {})
    data[DISCRETIZATION_MATRICES].setdefault(keyword, {})
    return data
