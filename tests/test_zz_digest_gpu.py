"""Parity AT THE BENCHMARKED SIZES against the reference: every output matrix of ``pb.Mpfa / pb.Mpsa / pb.Biot``
on Cartesian 32^3 (BASELINE config[0]), structured tetrahedra 12^3 x 6 and 16^3 x 6, and Biot 16^3 is compared
with a digest the unmodified reference produced on the very same arrays (tools/make_digests.py:
M @ x and |M| @ 1 on strided rows, bilinear forms over all entries, 200 sampled rows entrywise).  Tolerance 1e-10 relative (north star)."""
import os

import numpy as np
import pytest

import porepy_b200 as pb
from cases import DIGEST_CASES, digest_errors, digest_grid, digest_params, flatten
from golden_io import GOLDEN_DIR

TOL = 1e-10


def load_digest(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in d.files:
        if k.startswith("D__"):
            _, key, field = k.split("__")
            out.setdefault(key, {})[field] = d[k]
    return out


def discretize(name):
    kind, dims, what = DIGEST_CASES[name]
    g = digest_grid(kind, dims)
    k, bc, C, vbc, alpha = digest_params(g)
    if what == "mpfa":
        data = pb.initialize_data({}, "flow", {"second_order_tensor": k, "bc": bc})
        pb.Mpfa("flow").discretize(g, data)
        return data[pb.DISCRETIZATION_MATRICES]["flow"]
    prm = {"fourth_order_tensor": C, "bc": vbc}
    if what == "biot":
        prm["scalar_vector_mappings"] = {"flow": alpha}
    data = pb.initialize_data({}, "mech", prm)
    (pb.Biot if what == "biot" else pb.Mpsa)("mech").discretize(g, data)
    return flatten(data[pb.DISCRETIZATION_MATRICES]["mech"])


def check(name, mats):
    dig = load_digest(name)
    assert set(dig) == set(mats), (sorted(dig), sorted(mats))
    worst = 0.0
    for key, dg in dig.items():
        errs = digest_errors(dg, mats[key])
        assert max(errs) < TOL, (name, key, errs)
        worst = max(worst, *errs)
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DIGEST_CASES))
def test_benchmark_size_parity_with_the_reference(name):
    if not os.path.exists(os.path.join(GOLDEN_DIR, name + ".npz")):
        pytest.fail(f"fixture {name}.npz missing: run tools/make_digests.py in the build container")
    worst = check(name, discretize(name))
    print(f"{name}: worst relative digest error {worst:.2e}")
