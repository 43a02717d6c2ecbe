"""Import the read-only PorePy reference (/root/reference/src) in this container.

Test infrastructure only (used by tools/make_golden.py and the optional
"reference present" tests).  Seven third-party modules the hot path never
touches are absent from the image; they are stubbed before ``import porepy``
(SURVEY.md §8c).  Nothing here runs on the GPU box (``/root/reference`` does
not exist there).
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REF_SRC = "/root/reference/src"

_STUBS = [
    "meshio", "gmsh", "shapely", "shapely.geometry", "shapely.speedups",
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.tri",
    "matplotlib.patches", "matplotlib.figure", "matplotlib.axes",
    "matplotlib.ticker", "matplotlib.lines", "matplotlib.collections",
    "matplotlib.cm", "matplotlib.animation", "mpl_toolkits",
    "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
    "mpl_toolkits.axes_grid1", "deepdiff", "seaborn", "future",
]


class _Stub(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return MagicMock()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "porepy"))


def load_porepy():
    """Return the reference ``porepy`` module, or raise ImportError."""
    if "porepy" in sys.modules:
        return sys.modules["porepy"]
    if not reference_available():
        raise ImportError("reference tree /root/reference/src not present")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import porepy  # noqa: E402

    return porepy
