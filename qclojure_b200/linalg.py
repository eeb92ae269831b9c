"""P2 — the part of the reference's pluggable complex-linear-algebra backend that the simulation path
uses (src/org/soulspace/qclojure/domain/math/protocols.clj `MatrixAlgebra`; facade functions
domain/math/complex_linear_algebra.clj:228-446), computed on the GPU through `qcb_la_*`.

Method names mirror the protocol (kebab-case -> snake_case).  Inputs/outputs are NumPy complex128
arrays (the Clojure shim converts Vec2 vectors).  Conventions follow the reference's default backend
(math/fastmath/complex_linear_algebra.clj): `inner_product` conjugates its FIRST argument,
`outer_product` conjugates its SECOND.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _c(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


class B200ComplexBackend:
    def __init__(self, device: int = -1):
        self._sv = L.StateVector(1, device=device)
        self._lib, self._h = self._sv._lib, self._sv._h

    def close(self):
        self._sv.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        L.check(rc, self._h)

    def matrix_multiply(self, A, B):
        A, B = _c(A), _c(B)
        m, k = A.shape
        k2, n = B.shape
        assert k == k2
        out = np.empty((m, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_matmul(self._h, A.ctypes.data, B.ctypes.data, m, k, n, out.ctypes.data))
        return out

    def matrix_vector_product(self, A, x):
        A, x = _c(A), _c(x)
        out = np.empty(A.shape[0], dtype=np.complex128)
        self._ck(self._lib.qcb_la_matvec(self._h, A.ctypes.data, x.ctypes.data, A.shape[0], A.shape[1], out.ctypes.data))
        return out

    def kronecker_product(self, A, B):
        A, B = _c(A), _c(B)
        out = np.empty((A.shape[0] * B.shape[0], A.shape[1] * B.shape[1]), dtype=np.complex128)
        self._ck(self._lib.qcb_la_kron(self._h, A.ctypes.data, A.shape[0], A.shape[1], B.ctypes.data, B.shape[0], B.shape[1], out.ctypes.data))
        return out

    def inner_product(self, x, y):
        x, y = _c(x), _c(y)
        out = np.empty(1, dtype=np.complex128)
        self._ck(self._lib.qcb_la_inner(self._h, x.ctypes.data, y.ctypes.data, x.shape[0], out.ctypes.data))
        return complex(out[0])

    def outer_product(self, x, y):
        x, y = _c(x), _c(y)
        out = np.empty((x.shape[0], y.shape[0]), dtype=np.complex128)
        self._ck(self._lib.qcb_la_outer(self._h, x.ctypes.data, y.ctypes.data, x.shape[0], y.shape[0], out.ctypes.data))
        return out

    def trace(self, A):
        A = _c(A)
        out = np.empty(1, dtype=np.complex128)
        self._ck(self._lib.qcb_la_trace(self._h, A.ctypes.data, A.shape[0], out.ctypes.data))
        return complex(out[0])

    def norm2(self, x):
        x = _c(x).reshape(-1)
        v = C.c_double()
        self._ck(self._lib.qcb_la_norm2(self._h, x.ctypes.data, x.shape[0], C.byref(v)))
        return float(v.value)

    def _axpby(self, alpha, x, beta, y):
        x = _c(x)
        shape = x.shape
        xf = x.reshape(-1)
        yf = _c(y).reshape(-1) if y is not None else None
        out = np.empty(xf.shape[0], dtype=np.complex128)
        a = np.array([alpha], dtype=np.complex128)
        b = np.array([beta], dtype=np.complex128)
        self._ck(self._lib.qcb_la_axpby(self._h, a.ctypes.data, xf.ctypes.data, b.ctypes.data,
                                        yf.ctypes.data if yf is not None else None, xf.shape[0], out.ctypes.data))
        return out.reshape(shape)

    def add(self, A, B):
        return self._axpby(1.0, A, 1.0, B)

    def subtract(self, A, B):
        return self._axpby(1.0, A, -1.0, B)

    def scale(self, A, alpha):
        return self._axpby(alpha, A, 0.0, None)
