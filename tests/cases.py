"""Shared helpers: golden-case -> inputs for oracle / emulation / GPU product."""
from __future__ import annotations

import numpy as np

from golden_io import load_case, rel_err  # noqa: F401


def scalar_codes(bc, nf):
    internal = np.asarray(bc.is_internal, bool)
    codes = np.zeros(nf, np.uint8)
    codes[np.asarray(bc.is_neu, bool) | internal] = 2
    codes[np.asarray(bc.is_dir, bool) & ~internal] = 1
    codes[np.asarray(bc.is_rob, bool) & ~internal] = 3
    return codes


def vector_codes(bc, nd, nf):
    codes = np.zeros((nd, nf), np.uint8)
    codes[np.asarray(bc.is_neu, bool)[:nd]] = 2
    codes[np.asarray(bc.is_dir, bool)[:nd]] = 1
    codes[np.asarray(bc.is_rob, bool)[:nd]] = 3
    return codes


def flatten(out: dict) -> dict:
    """{'key': m, 'key2': {'a': m}} -> {'key': m, 'key2:a': m}"""
    flat = {}
    for k, v in out.items():
        if isinstance(v, dict):
            for ak, m in v.items():
                flat[f"{k}:{ak}"] = m
        else:
            flat[k] = v
    return flat


def max_rel_err(ref: dict, got: dict) -> tuple[float, str]:
    worst, wk = 0.0, ""
    for k, m in flatten(got).items():
        e = rel_err(ref[k], m)
        if e > worst:
            worst, wk = e, k
    return worst, wk
