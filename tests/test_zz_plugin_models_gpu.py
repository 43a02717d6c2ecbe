"""The PorePy model tests of tests/test_porepy_plugin.py with the REAL device plan: whole ``pp.SinglePhaseFlow`` (with
a fracture plane), ``pp.Poromechanics``, ``MomentumBalance``, ``Thermoporomechanics`` and ``MassAndEnergyBalance``
runs of the unmodified reference (``oracle/_ref``, placed by oracle/make_ref.sh and shipped to the GPU box)
discretizing through the CUDA kernels, same solution vectors as the stock models, zero reference fall-backs."""
import pytest

from test_porepy_plugin import (pp, reference_available,  # noqa: F401  (pp is a fixture)
                                test_lower_dimensional_grids_run_on_the_b200_path,  # noqa: F401
                                test_other_model_families, test_poromechanics_model,  # noqa: F401
                                test_refusals_propagate_unless_fallback_is_opted_in,  # noqa: F401
                                test_single_phase_flow_model_with_a_fracture)  # noqa: F401

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(), reason="oracle/_ref not present (run oracle/make_ref.sh)")]


@pytest.fixture()
def emu_plan():
    """Nothing to replace: the real ``DevicePlan`` / ``FaceGrid`` run."""
    from porepy_b200 import _lib
    _lib.require_gpu()
    yield
