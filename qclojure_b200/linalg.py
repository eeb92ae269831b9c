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
    """`host_only=True` skips the device handle: the decompositions / matrix functions / predicates below run on the
    host inside libqcb200 (qcb_la_* with a NULL handle); the MatrixAlgebra products need the GPU handle."""

    def __init__(self, device: int = -1, host_only: bool = False):
        if host_only:
            self._sv, self._lib, self._h = None, L.load(), None
        else:
            self._sv = L.StateVector(1, device=device)
            self._lib, self._h = self._sv._lib, self._sv._h

    def close(self):
        if self._sv is not None:
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

    # ------------------------------------------------------------------ host-side protocol methods (la_host.cpp)
    def _sq(self, A):
        A = _c(A)
        assert A.ndim == 2 and A.shape[0] == A.shape[1], "square matrix expected"
        return A, A.shape[0]

    def hadamard_product(self, A, B):
        A, B = _c(A), _c(B)
        out = np.empty_like(A)
        self._ck(self._lib.qcb_la_hadamard(self._h, A.ctypes.data, B.ctypes.data, A.size, out.ctypes.data))
        return out

    def transpose(self, A, conjugate: bool = False):
        A = _c(A)
        out = np.empty((A.shape[1], A.shape[0]), dtype=np.complex128)
        self._ck(self._lib.qcb_la_transpose(self._h, A.ctypes.data, A.shape[0], A.shape[1], int(conjugate), out.ctypes.data))
        return out

    def conjugate_transpose(self, A):
        return self.transpose(A, True)

    def negate(self, A):
        return -_c(A)

    def shape(self, A):
        return list(np.shape(A))

    def solve_linear_system(self, A, b):
        A, n = self._sq(A)
        b = _c(b)
        B = b.reshape(n, -1)
        out = np.empty_like(B)
        self._ck(self._lib.qcb_la_solve(self._h, A.ctypes.data, np.ascontiguousarray(B).ctypes.data, n, B.shape[1], out.ctypes.data))
        return out.reshape(b.shape)

    def inverse(self, A):
        A, n = self._sq(A)
        out = np.empty_like(A)
        self._ck(self._lib.qcb_la_inverse(self._h, A.ctypes.data, n, out.ctypes.data))
        return out

    def _pred(self, fn, A, eps):
        A, n = self._sq(A)
        v = C.c_int32()
        self._ck(fn(self._h, A.ctypes.data, n, C.c_double(eps), C.byref(v)))
        return bool(v.value)

    def is_hermitian(self, A, eps: float = 1e-12):
        return self._pred(self._lib.qcb_la_is_hermitian, A, eps)

    def is_diagonal(self, A, eps: float = 1e-12):
        return self._pred(self._lib.qcb_la_is_diagonal, A, eps)

    def is_unitary(self, U, eps: float = 1e-12):
        return self._pred(self._lib.qcb_la_is_unitary, U, eps)

    def is_positive_semidefinite(self, A, eps: float = 1e-12):
        return self._pred(self._lib.qcb_la_is_positive_semidefinite, A, eps)

    def eigen_hermitian(self, A):
        A, n = self._sq(A)
        w = np.empty(n, dtype=np.float64)
        v = np.empty((n, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_eigen_hermitian(self._h, A.ctypes.data, n, w.ctypes.data, v.ctypes.data))
        return {"eigenvalues": w, "eigenvectors": [v[k].copy() for k in range(n)]}

    def eigen_general(self, A):
        A, n = self._sq(A)
        w = np.empty(n, dtype=np.complex128)
        v = np.empty((n, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_eigen_general(self._h, A.ctypes.data, n, w.ctypes.data, v.ctypes.data))
        return {"eigenvalues": w, "eigenvectors": [v[k].copy() for k in range(n)]}

    def svd(self, A):
        A = _c(A)
        m, n = A.shape
        U = np.empty((m, m), dtype=np.complex128)
        S = np.empty(min(m, n), dtype=np.float64)
        Vh = np.empty((n, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_svd(self._h, A.ctypes.data, m, n, U.ctypes.data, S.ctypes.data, Vh.ctypes.data))
        return {"U": U, "S": S, "V†": Vh}

    def lu_decomposition(self, A):
        A, n = self._sq(A)
        P, Lm, U = (np.empty((n, n), dtype=np.complex128) for _ in range(3))
        self._ck(self._lib.qcb_la_lu(self._h, A.ctypes.data, n, P.ctypes.data, Lm.ctypes.data, U.ctypes.data))
        return {"P": P, "L": Lm, "U": U}

    def qr_decomposition(self, A):
        A = _c(A)
        m, n = A.shape
        Q = np.empty((m, m), dtype=np.complex128)
        R = np.empty((m, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_qr(self._h, A.ctypes.data, m, n, Q.ctypes.data, R.ctypes.data))
        return {"Q": Q, "R": R}

    def cholesky_decomposition(self, A):
        A, n = self._sq(A)
        Lm = np.empty((n, n), dtype=np.complex128)
        self._ck(self._lib.qcb_la_cholesky(self._h, A.ctypes.data, n, Lm.ctypes.data))
        return {"L": Lm}

    def _fun(self, fn, A):
        A, n = self._sq(A)
        out = np.empty_like(A)
        self._ck(fn(self._h, A.ctypes.data, n, out.ctypes.data))
        return out

    def matrix_exp(self, A):
        return self._fun(self._lib.qcb_la_matrix_exp, A)

    def matrix_log(self, A):
        return self._fun(self._lib.qcb_la_matrix_log, A)

    def matrix_sqrt(self, A):
        return self._fun(self._lib.qcb_la_matrix_sqrt, A)

    def spectral_norm(self, A):
        A = _c(A)
        v = C.c_double()
        self._ck(self._lib.qcb_la_spectral_norm(self._h, A.ctypes.data, A.shape[0], A.shape[1], C.byref(v)))
        return float(v.value)

    def condition_number(self, A):
        A = _c(A)
        v = C.c_double()
        self._ck(self._lib.qcb_la_condition_number(self._h, A.ctypes.data, A.shape[0], A.shape[1], C.byref(v)))
        return float(v.value)
