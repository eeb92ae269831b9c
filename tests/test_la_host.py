"""P2 completeness: the decomposition / matrix-function / predicate methods of the reference's
math protocols (domain/math/protocols.clj:284-521) computed on the host inside libqcb200 (csrc/la_host.cpp), checked
against NumPy / SciPy and against the defining identities the reference's own tests use
(test/.../domain/math/complex_linear_algebra_test.clj:131-338, backend_test.clj:258-282): eigenvalues ascending,
A v = lambda v, U S V^H = A with S descending, P L U = A, Q R = A, L L^H = A, exp/log/sqrt round trips.  No GPU needed."""
import numpy as np
import pytest
import scipy.linalg as sla

from qclojure_b200 import _lib as L
from qclojure_b200.linalg import B200ComplexBackend

TOL = 1e-10


@pytest.fixture(scope="module")
def be():
    return B200ComplexBackend(host_only=True)


def _rand(rng, m, n=None):
    n = m if n is None else n
    return rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))


def _herm(rng, n):
    A = _rand(rng, n)
    return (A + A.conj().T) / 2


def _unitary(rng, n):
    return np.linalg.qr(_rand(rng, n))[0]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 8, 16])
def test_eigen_hermitian(be, n):
    rng = np.random.default_rng(n)
    A = _herm(rng, n)
    r = be.eigen_hermitian(A)
    w, vs = r["eigenvalues"], r["eigenvectors"]
    assert np.all(np.diff(w) >= -1e-12)
    assert np.max(np.abs(w - np.linalg.eigvalsh(A))) <= TOL
    V = np.array(vs).T
    assert np.max(np.abs(V.conj().T @ V - np.eye(n))) <= TOL
    for k in range(n):
        assert np.max(np.abs(A @ vs[k] - w[k] * vs[k])) <= TOL


def test_eigen_hermitian_reference_examples(be):
    """The matrices of the reference's REPL examples (fastmath/complex_linear_algebra.clj:1476-1488) and degenerate spectra."""
    r = be.eigen_hermitian([[1, 2], [2, 1]])
    assert np.allclose(r["eigenvalues"], [-1, 3], atol=TOL)
    r = be.eigen_hermitian([[3, 1], [1, 3]])
    assert np.allclose(r["eigenvalues"], [2, 4], atol=TOL)
    r = be.eigen_hermitian(np.diag([2.0, 2.0, 2.0]))
    assert np.allclose(r["eigenvalues"], [2, 2, 2], atol=TOL)
    Y = np.array([[0, -1j], [1j, 0]])
    r = be.eigen_hermitian(np.kron(Y, Y))
    assert np.allclose(r["eigenvalues"], [-1, -1, 1, 1], atol=TOL)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 12])
def test_eigen_general(be, n):
    rng = np.random.default_rng(100 + n)
    A = _rand(rng, n)
    r = be.eigen_general(A)
    w, vs = r["eigenvalues"], r["eigenvectors"]
    ref = np.linalg.eigvals(A)
    assert np.max(np.abs(np.sort_complex(w) - np.sort_complex(ref))) <= 1e-9
    assert all((w[k].real, w[k].imag) <= (w[k + 1].real + 1e-12, w[k + 1].imag + 1e300) for k in range(n - 1))
    for k in range(n):
        assert abs(np.linalg.norm(vs[k]) - 1.0) <= TOL
        assert np.max(np.abs(A @ vs[k] - w[k] * vs[k])) <= 1e-8
    U = _unitary(rng, n)                       # a unitary (normal, eigenvalues on the unit circle)
    w = be.eigen_general(U)["eigenvalues"]
    assert np.max(np.abs(np.abs(w) - 1.0)) <= 1e-9


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (4, 4), (6, 3), (3, 6), (8, 8), (5, 2)])
def test_svd(be, shape):
    rng = np.random.default_rng(shape[0] * 10 + shape[1])
    A = _rand(rng, *shape)
    r = be.svd(A)
    U, S, Vh = r["U"], r["S"], r["V†"]
    m, n = shape
    assert np.all(np.diff(S) <= 1e-12) and np.all(S >= 0)
    assert np.max(np.abs(S - np.linalg.svd(A, compute_uv=False))) <= TOL
    Sm = np.zeros((m, n), dtype=np.complex128)
    Sm[:len(S), :len(S)] = np.diag(S)
    assert np.max(np.abs(U @ Sm @ Vh - A)) <= TOL
    assert np.max(np.abs(U.conj().T @ U - np.eye(m))) <= TOL
    assert np.max(np.abs(Vh @ Vh.conj().T - np.eye(n))) <= TOL


def test_svd_rank_deficient(be):
    rng = np.random.default_rng(5)
    x, y = _rand(rng, 5, 1), _rand(rng, 1, 4)
    A = x @ y                                   # rank 1
    r = be.svd(A)
    assert np.max(np.abs(r["S"][1:])) <= 1e-9
    Sm = np.zeros((5, 4), dtype=np.complex128)
    Sm[:4, :4] = np.diag(r["S"])
    assert np.max(np.abs(r["U"] @ Sm @ r["V†"] - A)) <= TOL
    assert np.max(np.abs(r["U"].conj().T @ r["U"] - np.eye(5))) <= TOL


@pytest.mark.parametrize("n", [1, 2, 3, 6, 10])
def test_lu_qr_cholesky_inverse_solve(be, n):
    rng = np.random.default_rng(200 + n)
    A = _rand(rng, n)
    r = be.lu_decomposition(A)
    P, Lm, U = r["P"], r["L"], r["U"]
    assert np.max(np.abs(P @ Lm @ U - A)) <= TOL
    assert np.max(np.abs(np.triu(Lm, 1))) == 0 and np.max(np.abs(np.tril(U, -1))) == 0
    assert np.allclose(np.diag(Lm), 1) and np.allclose(P @ P.T, np.eye(n))
    q = be.qr_decomposition(A)
    assert np.max(np.abs(q["Q"] @ q["R"] - A)) <= TOL
    assert np.max(np.abs(q["Q"].conj().T @ q["Q"] - np.eye(n))) <= TOL
    assert np.max(np.abs(np.tril(q["R"], -1))) == 0
    H = A @ A.conj().T + 0.1 * np.eye(n)       # positive definite
    c = be.cholesky_decomposition(H)["L"]
    assert np.max(np.abs(c @ c.conj().T - H)) <= TOL and np.max(np.abs(np.triu(c, 1))) == 0
    assert np.max(np.abs(be.inverse(A) @ A - np.eye(n))) <= 1e-8
    b = _rand(rng, n, 1)[:, 0]
    assert np.max(np.abs(A @ be.solve_linear_system(A, b) - b)) <= 1e-9


def test_rectangular_qr(be):
    rng = np.random.default_rng(7)
    for shape in ((5, 3), (3, 5)):
        A = _rand(rng, *shape)
        q = be.qr_decomposition(A)
        assert np.max(np.abs(q["Q"] @ q["R"] - A)) <= TOL
        assert np.max(np.abs(np.tril(q["R"], -1))) == 0


def test_errors(be):
    with pytest.raises(L.QcbError):
        be.inverse([[1, 2], [2, 4]])
    with pytest.raises(L.QcbError):
        be.cholesky_decomposition([[1, 0], [0, -1]])
    with pytest.raises(L.QcbError):
        be.is_positive_semidefinite([[1, 2], [0, 1]])         # not Hermitian (reference throws)
    with pytest.raises(L.QcbError):
        be.matrix_log([[1, 0], [0, 0]])


def test_predicates(be):
    rng = np.random.default_rng(9)
    H, U = _herm(rng, 4), _unitary(rng, 4)
    assert be.is_hermitian(H) and not be.is_hermitian(U)
    assert be.is_unitary(U, 1e-10) and not be.is_unitary(H, 1e-10)
    assert be.is_diagonal(np.diag([1, 2j, 3])) and not be.is_diagonal(H)
    assert be.is_positive_semidefinite(H @ H.conj().T, 1e-10)
    assert not be.is_positive_semidefinite(H - 10 * np.eye(4), 1e-10)
    assert be.is_positive_semidefinite(np.zeros((3, 3)))
    A, B = _rand(rng, 3, 2), _rand(rng, 3, 2)
    assert np.allclose(be.hadamard_product(A, B), A * B)
    assert np.allclose(be.transpose(A), A.T) and np.allclose(be.conjugate_transpose(A), A.conj().T)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_matrix_functions(be, n):
    rng = np.random.default_rng(300 + n)
    A = _rand(rng, n)
    assert np.max(np.abs(be.matrix_exp(A) - sla.expm(A))) <= 1e-9 * max(1.0, np.linalg.norm(sla.expm(A)))
    H = _herm(rng, n)
    Ue = be.matrix_exp(-1j * H)                # time evolution operator: unitary
    assert np.max(np.abs(Ue.conj().T @ Ue - np.eye(n))) <= TOL
    assert np.max(np.abs(be.matrix_log(Ue) - sla.logm(Ue))) <= 1e-8 or np.max(np.abs(be.matrix_exp(be.matrix_log(Ue)) - Ue)) <= 1e-9
    P = A @ A.conj().T + np.eye(n)             # positive definite: principal sqrt / log are Hermitian
    R = be.matrix_sqrt(P)
    assert np.max(np.abs(R @ R - P)) <= 1e-9 * np.linalg.norm(P)
    assert np.max(np.abs(R - sla.sqrtm(P))) <= 1e-8 * np.linalg.norm(P)
    Lg = be.matrix_log(P)
    assert np.max(np.abs(be.matrix_exp(Lg) - P)) <= 1e-8 * np.linalg.norm(P)
    assert np.max(np.abs(Lg - sla.logm(P))) <= 1e-8
    G = A + 3 * n * np.eye(n)                  # general, spectrum away from the negative real axis
    R = be.matrix_sqrt(G)
    assert np.max(np.abs(R @ R - G)) <= 1e-9 * np.linalg.norm(G)
    assert np.max(np.abs(be.matrix_exp(be.matrix_log(G)) - G)) <= 1e-8 * np.linalg.norm(G)


def test_matrix_functions_on_gates(be):
    """exp(-i theta/2 X) = RX(theta) (domain/gate.clj:218-219); sqrt(X)^2 = X; log of a rotation."""
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    th = 0.73
    rx = np.array([[np.cos(th / 2), -1j * np.sin(th / 2)], [-1j * np.sin(th / 2), np.cos(th / 2)]])
    assert np.max(np.abs(be.matrix_exp(-0.5j * th * X) - rx)) <= TOL
    s = be.matrix_sqrt(X)
    assert np.max(np.abs(s @ s - X)) <= TOL
    assert np.max(np.abs(be.matrix_log(rx) - (-0.5j * th * X))) <= TOL


def test_norms(be):
    rng = np.random.default_rng(11)
    A = _rand(rng, 5, 3)
    s = np.linalg.svd(A, compute_uv=False)
    assert abs(be.spectral_norm(A) - s[0]) <= TOL
    assert abs(be.condition_number(A) - s[0] / s[-1]) <= 1e-9 * s[0] / s[-1]
    assert abs(be.spectral_norm(_unitary(rng, 4)) - 1.0) <= TOL
    assert be.condition_number([[1, 2], [2, 4]]) > 1e12
