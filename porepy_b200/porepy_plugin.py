"""Drop-in plugin for an installed PorePy: subclasses of ``pp.Mpfa / pp.Mpsa / pp.Biot`` and of
the AD wrappers ``pp.ad.MpfaAd / MpsaAd / BiotAd`` whose ``discretize`` runs on the GPU.

Why subclasses: ``MpfaAd.__init__`` hard-codes ``pp.Mpfa(keyword)`` (reference
src/porepy/numerics/ad/discretizations.py:192-206) and the model mixins test
``isinstance(x, pp.ad.MpfaAd)`` / ``(MpsaAd, BiotAd)`` (models/constitutive_laws.py:1341,1457,
2903), so the cores must be ``pp.Mpfa`` etc. and the wrappers ``pp.ad.MpfaAd`` etc.
``uniquify_discretization_list`` keys on ``(discr._discr.__class__, keyword)``
(numerics/ad/ad_utils.py:244-276), which stays unique for the subclasses.

Use in a model (the mixin override of constitutive_laws.py:1078,3003,3506)::

    import porepy as pp
    from porepy_b200.porepy_plugin import plugin
    b200 = plugin(pp)

    class B200Flow(pp.SinglePhaseFlow):
        def darcy_flux_discretization(self, subdomains):
            return b200.MpfaAd(self.darcy_keyword, subdomains)

    class B200Poromechanics(b200.ModelMixin, pp.Poromechanics):   # all hooks at once
        pass

Scope: the GPU path covers the top-dimensional subdomains (3-D grids, 2-D grids lying in the
xy-plane) and, for the flux discretization, 2-D fracture planes embedded in 3-D.  Intersection
grids (1-D lines, 0-D points) are handed to the reference's own implementation (its TPFA
fallback, mpfa.py:690-712), exactly as ``pp.Mpfa`` would.
"""
from __future__ import annotations

import logging
from types import SimpleNamespace

import numpy as np

from . import fv

logger = logging.getLogger(__name__)


def _gpu_scope(sd, flow: bool = False) -> bool:
    """3-D grids; 2-D grids in a plane z = const; for the flux discretization also 2-D fracture
    planes embedded in 3-D (rotated into their plane on the host, fv.plane_frame)."""
    if sd.dim == 3:
        return True
    if sd.dim == 2:
        if flow:
            return True
        z = np.asarray(sd.nodes)[2]
        return bool(np.ptp(z) <= 1e-12 * max(1.0, float(np.abs(sd.nodes).max())))
    return False


def plugin(pp) -> SimpleNamespace:
    """Build the plugin classes against the given ``porepy`` module."""

    class Mpfa(fv.Mpfa, pp.Mpfa):
        def __init__(self, keyword: str) -> None:
            pp.Mpfa.__init__(self, keyword)
            fv.Mpfa.__init__(self, keyword)

        def discretize(self, sd, data) -> None:
            if _gpu_scope(sd, flow=True) and not hasattr(sd, "periodic_face_map"):
                fv.Mpfa.discretize(self, sd, data)
            else:
                logger.info("B200 Mpfa: %s-d subdomain outside the GPU scope -> reference path", sd.dim)
                pp.Mpfa.discretize(self, sd, data)

        def assemble_matrix_rhs(self, sd, data):
            return pp.Mpfa.assemble_matrix_rhs(self, sd, data)

        def update_discretization(self, sd, data) -> None:
            self.discretize(sd, data)

    class Mpsa(fv.Mpsa, pp.Mpsa):
        def __init__(self, keyword: str) -> None:
            pp.Mpsa.__init__(self, keyword)
            fv.Mpsa.__init__(self, keyword)

        def discretize(self, sd, data) -> None:
            if _gpu_scope(sd):
                fv.Mpsa.discretize(self, sd, data)
            else:
                pp.Mpsa.discretize(self, sd, data)

        def assemble_matrix_rhs(self, sd, data):
            return pp.Mpsa.assemble_matrix_rhs(self, sd, data)

        def update_discretization(self, sd, data) -> None:
            self.discretize(sd, data)

    class Biot(fv.Biot, pp.Biot):
        def __init__(self, keyword: str = "mechanics") -> None:
            pp.Biot.__init__(self, keyword)
            fv.Biot.__init__(self, keyword)

        def discretize(self, sd, data) -> None:
            if _gpu_scope(sd):
                fv.Biot.discretize(self, sd, data)
            else:
                pp.Biot.discretize(self, sd, data)

        def update_discretization(self, sd, data) -> None:
            self.discretize(sd, data)

    def _rewrap(obj, discr, subdomains, coupling_terms=None):
        obj._discretization = discr
        if coupling_terms is None:
            pp.ad.wrap_discretization(obj, discr, subdomains=subdomains)
        else:
            pp.ad.wrap_discretization(obj=obj, discr=discr, subdomains=subdomains,
                                      coupling_terms=coupling_terms)

    class MpfaAd(pp.ad.MpfaAd):
        def __init__(self, keyword, subdomains):
            super().__init__(keyword, subdomains)
            _rewrap(self, Mpfa(keyword), subdomains)

    class MpsaAd(pp.ad.MpsaAd):
        def __init__(self, keyword, subdomains):
            super().__init__(keyword, subdomains)
            _rewrap(self, Mpsa(keyword), subdomains)

    class BiotAd(pp.ad.BiotAd):
        def __init__(self, keyword, subdomains):
            super().__init__(keyword, subdomains)
            _rewrap(self, Biot(keyword), subdomains,
                    ["displacement_divergence", "bound_displacement_divergence", "scalar_gradient",
                     "bound_pressure", "consistency"])

    class ModelMixin:
        """Put FIRST among the bases of a PorePy model class to route its flux / stress
        discretizations through the GPU classes::

            class Model(b200.ModelMixin, Geometry, BoundaryConditions, pp.Poromechanics): ...

        Overrides the constitutive-law hooks (models/constitutive_laws.py:1078, 2425, 3003, 3506)
        and relaxes the exact-type check of ``add_nonlinear_diffusive_flux_discretization``
        (models/solution_strategy.py:505-524: ``type(x) in [pp.Mpfa, pp.Tpfa]``) to ``isinstance``,
        which the plugin's subclasses satisfy."""

        def darcy_flux_discretization(self, subdomains):
            return MpfaAd(self.darcy_keyword, subdomains)

        def fourier_flux_discretization(self, subdomains):
            return MpfaAd(self.fourier_keyword, subdomains)

        def stress_discretization(self, subdomains):
            stock = super().stress_discretization(subdomains)
            cls = BiotAd if isinstance(stock, pp.ad.BiotAd) else MpsaAd
            return cls(self.stress_keyword, subdomains)

        def add_nonlinear_diffusive_flux_discretization(self, discretization) -> None:
            if not isinstance(discretization._discr, (pp.Mpfa, pp.Tpfa)):
                raise TypeError(f"Expecting an Mpfa or Tpfa discretization, got {type(discretization._discr)}")
            if discretization not in self._nonlinear_diffusive_flux_discretizations:
                self._nonlinear_diffusive_flux_discretizations.append(discretization)

    return SimpleNamespace(Mpfa=Mpfa, Mpsa=Mpsa, Biot=Biot, MpfaAd=MpfaAd, MpsaAd=MpsaAd, BiotAd=BiotAd,
                           ModelMixin=ModelMixin)
