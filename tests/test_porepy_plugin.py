"""Wiring of the PorePy plugin classes (porepy_b200/porepy_plugin.py) against the reference,
when the reference tree is importable (build container only; skipped on the GPU box)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_loader import load_porepy, reference_available  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def pp():
    return load_porepy()


def test_plugin_classes_are_reference_subclasses(pp):
    from porepy_b200.porepy_plugin import plugin
    b = plugin(pp)
    g = pp.CartGrid([2, 2, 2])
    g.compute_geometry()
    ad = b.MpfaAd("flow", [g])
    assert isinstance(ad, pp.ad.MpfaAd) and isinstance(ad._discretization, pp.Mpfa)
    assert type(ad._discretization).discretize is b.Mpfa.discretize
    assert str(ad.flux()) == "Mpfa(flow).flux"
    bad = b.BiotAd("mech", [g])
    assert isinstance(bad, pp.ad.BiotAd) and isinstance(bad._discretization, pp.Biot)
    for term in ("displacement_divergence", "bound_displacement_divergence", "scalar_gradient",
                 "bound_pressure", "consistency"):
        assert callable(getattr(bad, term))
    sad = b.MpsaAd("mech", [g])
    assert isinstance(sad, pp.ad.MpsaAd) and sad._discretization.ndof(g) == 24
    # same matrix keys as the reference cores
    for cls_b, cls_r in ((b.Mpfa, pp.Mpfa), (b.Mpsa, pp.Mpsa), (b.Biot, pp.Biot)):
        kb = {k: v for k, v in vars(cls_b("kw")).items() if k.endswith("_matrix_key")}
        kr = {k: v for k, v in vars(cls_r("kw")).items() if k.endswith("_matrix_key")}
        assert kb == kr


def test_lower_dimensional_grids_take_the_reference_path(pp):
    from porepy_b200.porepy_plugin import plugin
    b = plugin(pp)
    g1 = pp.CartGrid([4])
    g1.compute_geometry()
    d1 = pp.initialize_data({}, "flow", {"second_order_tensor": pp.SecondOrderTensor(np.ones(4)),
                                         "bc": pp.BoundaryCondition(g1)})
    b.Mpfa("flow").discretize(g1, d1)
    assert d1[pp.DISCRETIZATION_MATRICES]["flow"]["flux"].shape == (5, 4)


def test_parameter_mirrors_match_reference(pp):
    import porepy_b200 as pb
    rng = np.random.default_rng(0)
    mu, lam = rng.random(5) + 1, rng.random(5)
    assert np.allclose(pp.FourthOrderTensor(mu, lam).values, pb.FourthOrderTensor(mu, lam).values)
    a = [1 + rng.random(5) for _ in range(3)] + [0.2 * rng.random(5) for _ in range(3)]
    assert np.allclose(pp.SecondOrderTensor(*a).values, pb.SecondOrderTensor(*a).values)
    assert pb.PARAMETERS == pp.PARAMETERS and pb.DISCRETIZATION_MATRICES == pp.DISCRETIZATION_MATRICES
    g = pp.CartGrid([2, 2, 2])
    g.compute_geometry()
    bf = g.get_all_boundary_faces()
    r = pp.BoundaryCondition(g, bf[:5], ["dir"] * 5)
    m = pb.BoundaryCondition(g, bf[:5], ["dir"] * 5)
    assert np.array_equal(r.is_dir, m.is_dir) and np.array_equal(r.is_neu, m.is_neu)
