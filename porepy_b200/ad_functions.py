"""Elementwise functions of ``DeviceAdArray`` -- the counterpart of the reference's ``pp.ad.functions``
(src/porepy/numerics/ad/functions.py:60-490) and of ``AdArray.__pow__`` (numerics/ad/forward_mode.py:341-406): value by a
torch kernel on the device, Jacobian by a row scaling of the operand's Jacobian (``_diagvec_mul_jac``, forward_mode.py:613-616)
or, for ``maximum`` / ``l2_norm``, a row selection / a small SpGEMM.  What the contact-mechanics and friction laws of the
models are written with (``maximum``, ``l2_norm``, ``characteristic_function``, ``heaviside``).

Every function also accepts plain tensors (the reference accepts ndarrays): operators evaluated at a previous iterate.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

from . import ad
from .ad import DeviceAdArray


def _t():
    import torch
    return torch


def _unary(var, f, df):
    if isinstance(var, DeviceAdArray):
        return DeviceAdArray(f(var.val), var.jac.scaled(df(var.val).contiguous()))
    return f(ad.device_vector(var))


def exp(var):
    return _unary(var, lambda v: v.exp(), lambda v: v.exp())


def log(var):
    return _unary(var, lambda v: v.log(), lambda v: 1.0 / v)


def abs(var):  # noqa: A001 - the reference's name
    return _unary(var, lambda v: v.abs(), lambda v: v.sign())


def sqrt(var):
    return _unary(var, lambda v: v.sqrt(), lambda v: 0.5 / v.sqrt())


def sin(var):
    return _unary(var, lambda v: v.sin(), lambda v: v.cos())


def cos(var):
    return _unary(var, lambda v: v.cos(), lambda v: -v.sin())


def tanh(var):
    return _unary(var, lambda v: v.tanh(), lambda v: 1.0 - v.tanh() ** 2)


def power(var, exponent: float):
    """``var ** exponent`` for a scalar exponent (forward_mode.py:357-360)."""
    e = float(exponent)
    return _unary(var, lambda v: v ** e, lambda v: e * v ** (e - 1.0))


def heaviside(zerovalue: float, var):
    """0 / ``zerovalue`` / 1 for negative / zero / positive entries, zero Jacobian (functions.py:289-314)."""
    torch = _t()

    def h(v):
        return torch.heaviside(v, torch.as_tensor(float(zerovalue), dtype=v.dtype, device=v.device))
    if isinstance(var, DeviceAdArray):
        return DeviceAdArray(h(var.val), var.jac * 0.0)
    return h(ad.device_vector(var))


def heaviside_smooth(var, eps: float = 1e-3):
    """(1 + (2 / pi) arctan(x / eps)) / 2 (functions.py:317-343)."""
    return _unary(var, lambda v: 0.5 * (1.0 + (2.0 / np.pi) * (v / eps).atan()),
                  lambda v: (eps / np.pi) / (eps ** 2 + v ** 2))


def characteristic_function(tol: float, var):
    """1 where ``|var| <= tol`` (``np.isclose(var, 0, atol=tol)``), zero Jacobian (functions.py:463-490)."""
    def chi(v):
        return (v.abs() <= tol).to(v.dtype)
    if isinstance(var, DeviceAdArray):
        return DeviceAdArray(chi(var.val), var.jac * 0.0)
    return chi(ad.device_vector(var))


def maximum(var_0, var_1):
    """Elementwise maximum; at equality the FIRST argument wins, value and Jacobian row (functions.py:360-460).
    Arguments: ``DeviceAdArray``, tensors / arrays (zero Jacobian) or scalars (broadcast)."""
    torch = _t()
    ads = [v for v in (var_0, var_1) if isinstance(v, DeviceAdArray)]
    if not ads:
        return torch.maximum(*(ad.device_vector(np.atleast_1d(v)) if not torch.is_tensor(v) else v for v in (var_0, var_1)))
    like = ads[0]

    def parts(v):
        if isinstance(v, DeviceAdArray):
            return v.val, v.jac
        if isinstance(v, (int, float, np.floating, np.integer)):
            return torch.full_like(like.val, float(v)), None
        return ad.device_vector(v), None
    (v0, j0), (v1, j1) = parts(var_0), parts(var_1)
    second = v1 > v0                                   # rows taken from the second argument
    val = torch.where(second, v1, v0)
    w1 = second.to(val.dtype).contiguous()
    w0 = (1.0 - w1).contiguous()
    zero = like.jac * 0.0
    j0 = zero if j0 is None else j0
    j1 = zero if j1 is None else j1
    return DeviceAdArray(val, j0.scaled(w0).axpby(1.0, j1.scaled(w1), 1.0))


def l2_norm(dim: int, var):
    """Euclidean norm of consecutive ``dim``-vectors ``[u0, v0, w0, u1, ...]`` (functions.py:90-142): the Jacobian is
    ``N @ var.jac`` with one row of unit-vector components per vector; vectors shorter than 1e-12 get entries 1 as in the
    reference."""
    torch = _t()
    if not isinstance(var, DeviceAdArray):
        v = ad.device_vector(var)
        return torch.linalg.vector_norm(v.reshape(-1, dim), dim=1)
    if dim == 1:
        return abs(var)
    n = var.val.numel()
    if n % dim:
        raise ValueError("the array does not hold whole vectors")
    size = n // dim
    resh = var.val.reshape(size, dim)
    vals = torch.linalg.vector_norm(resh, dim=1)
    unit = torch.where((vals > 1e-12)[:, None], resh / vals.clamp_min(1e-300)[:, None], torch.ones_like(resh))
    # structure on the host (index arithmetic only), values on the device through a column scaling of the 0/1 pattern
    rows = np.repeat(np.arange(size), dim)
    pattern = ad.as_device_csr(sps.csr_matrix((np.ones(n), (rows, np.arange(n))), shape=(size, n)))
    norm_jac = pattern.scaled(unit.reshape(-1).contiguous(), by_cols=True)
    return DeviceAdArray(vals, norm_jac @ var.jac)
