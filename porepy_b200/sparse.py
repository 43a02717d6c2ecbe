"""Device-resident CSR matrix with the SpMV the Newton residual / Krylov loop needs.

Replaces the scipy ``M @ val`` of ``AdArray.__rmatmul__`` (reference
src/porepy/numerics/ad/forward_mode.py:565-595) for matrices kept in HBM."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sps

from . import _lib


class DeviceCsr:
    def __init__(self, a):
        lib = _lib.load()
        _lib.require_gpu()
        a = sps.csr_matrix(a)
        a.sort_indices()
        if a.nnz >= 2**31:
            raise ValueError("matrix too large for int32 indices")
        self.shape = a.shape
        self.nnz = int(a.nnz)
        ip = a.indptr.astype(np.int32)
        ix = a.indices.astype(np.int32)
        da = np.ascontiguousarray(a.data, np.float64)
        h = C.c_void_p()
        _lib.check(lib.pb_csr_create(a.shape[0], a.shape[1], a.nnz, _lib.ptr(ip, _lib._i32p),
                                     _lib.ptr(ix, _lib._i32p), _lib.ptr(da, _lib._f64p), C.byref(h)))
        self.h = h
        self.lib = lib

    @classmethod
    def from_handle(cls, h) -> "DeviceCsr":
        """Wrap a matrix that was assembled on the device (e.g. ``DevicePlan.mpfa_system``)."""
        lib = _lib.load()
        self = cls.__new__(cls)
        nr, ncl, nz = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.pb_csr_shape(h, C.byref(nr), C.byref(ncl), C.byref(nz)))
        self.shape, self.nnz, self.h, self.lib = (nr.value, ncl.value), nz.value, h, lib
        return self

    def truncate_rows(self, nrows: int) -> "DeviceCsr":
        """Keep the first ``nrows`` rows (in place; the own rows of a shard's system come first)."""
        _lib.check(self.lib.pb_csr_truncate_rows(self.h, int(nrows)))
        self.shape = (int(nrows), self.shape[1])
        return self

    def to_scipy(self) -> sps.csr_matrix:
        ip = np.empty(self.shape[0] + 1, np.int32)
        ix = np.empty(max(self.nnz, 1), np.int32)
        da = np.empty(max(self.nnz, 1), np.float64)
        _lib.check(self.lib.pb_csr_download(self.h, _lib.ptr(ip, _lib._i32p), _lib.ptr(ix, _lib._i32p),
                                            _lib.ptr(da, _lib._f64p)))
        nnz = int(ip[-1])   # after truncate_rows the row-pointer prefix addresses fewer entries
        return sps.csr_matrix((da[:nnz], ix[:nnz], ip), shape=self.shape)

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value:
            try:
                self.lib.pb_csr_destroy(h)
            except Exception:
                pass
            self.h = None

    def __matmul__(self, x):
        x = _lib.f64(x)
        if x.shape != (self.shape[1],):
            raise ValueError("dimension mismatch")
        y = np.empty(self.shape[0])
        _lib.check(self.lib.pb_csr_spmv(self.h, _lib.ptr(x, _lib._f64p), _lib.ptr(y, _lib._f64p)))
        return y

    def spmv_device(self, x_ptr: int, y_ptr: int, stream: int = 0) -> None:
        """y = A x on raw device pointers (e.g. ``torch.Tensor.data_ptr()``)."""
        _lib.check(self.lib.pb_csr_spmv_dev(self.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), stream))

    def bench(self, reps: int = 50) -> float:
        """Mean device milliseconds per SpMV (CUDA events on the launching stream)."""
        ms = C.c_float()
        _lib.check(self.lib.pb_csr_spmv_bench(self.h, reps, C.byref(ms)))
        return float(ms.value)

    def algorithmic_bytes(self) -> int:
        """12 B per non-zero + 20 B per row."""
        return 12 * self.nnz + 20 * self.shape[0]
