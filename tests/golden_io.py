"""Load the golden fixtures written by tools/make_golden.py (reference outputs)."""
from __future__ import annotations

import glob
import os
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sps

from porepy_b200.grid import Grid

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(prefix: str):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_case(name: str):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g = Grid.from_arrays(d)
    mats = {}
    for k in d:
        if k.startswith("M__") and k.endswith("__data"):
            key = k[3:-6]
            mats[key] = sps.csr_matrix((d[f"M__{key}__data"], d[f"M__{key}__indices"],
                                        d[f"M__{key}__indptr"]), shape=tuple(d[f"M__{key}__shape"]))
    bc = None
    if "bc_is_dir" in d:               # fixtures of grid-only routines (geometry, DifferentiableTpfa) carry no bc
        bc = SimpleNamespace(is_dir=d["bc_is_dir"], is_neu=d["bc_is_neu"], is_rob=d["bc_is_rob"],
                             is_internal=d["bc_is_internal"], robin_weight=d["bc_robin_weight"],
                             bc_type="vectorial" if d["bc_is_dir"].ndim == 2 else "scalar",
                             num_faces=g.num_faces)
        if "bc_basis" in d:
            bc.basis = d["bc_basis"]
    alpha = {k[7:]: d[k] for k in d if k.startswith("alpha__")}
    return SimpleNamespace(name=name, kind=str(d["kind"]), g=g, bc=bc, mats=mats, raw=d,
                           eta=float(d["eta"]) if "eta" in d else 0.0, alpha=alpha)


def rel_err(ref, got) -> float:
    """max |ref - got| / max |ref| on the densified difference (pattern agnostic, the
    comparator of applications/test_utils/arrays.py:49-74 made relative)."""
    ref = sps.csr_matrix(ref)
    got = sps.csr_matrix(got)
    scale = abs(ref).max() if ref.nnz else 1.0
    diff = abs(ref - got)
    return float(diff.max() / scale) if diff.nnz else 0.0
