"""The digest comparison itself, on CPU: the smallest benchmark-size fixture against the host build of the
node routines (the GPU test tests/test_zz_digest_gpu.py runs all of them through the CUDA path)."""
import numpy as np
import pytest

from emu_binding import EmuBackedPlan
from porepy_b200 import fv
from test_zz_digest_gpu import check, discretize, load_digest


def test_smallest_digest_against_the_host_build(monkeypatch):
    monkeypatch.setattr(fv, "DevicePlan", EmuBackedPlan)
    mats = discretize("digest_mpfa_tet12")
    assert check("digest_mpfa_tet12", mats) < 1e-10
    # the comparison is sensitive: one perturbed coefficient is seen
    dig = load_digest("digest_mpfa_tet12")
    m = mats["flux"].copy()
    rows = dig["flux"]["rows"]
    best = max(rows, key=lambda r: np.abs(m.data[m.indptr[r]:m.indptr[r + 1]]).max(initial=0.0))
    lo, hi = m.indptr[best], m.indptr[best + 1]
    m.data[lo + int(np.argmax(np.abs(m.data[lo:hi])))] *= 1 + 1e-8   # one coefficient of a sampled row, 1e-8 relative
    from cases import digest_errors
    assert max(digest_errors(dig["flux"], m)) > 1e-10
