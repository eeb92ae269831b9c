// la_host.cpp — the remaining methods of the reference's pluggable complex-linear-algebra backend:
// MatrixAlgebra (solve-linear-system, inverse, predicates, hadamard-product, transpose), MatrixDecompositions,
// MatrixFunctions and MatrixAnalysis of src/org/soulspace/qclojure/domain/math/protocols.clj:81-521.
//
// These act on the SMALL dense matrices of the facade (gate matrices, observables, density matrices of a few qubits):
// they are plain host C++ (the survey: "small dense matrices -> host C++ (Jacobi/QR) is adequate"), not GPU kernels, and
// are not on the measured hot path.  Conventions follow the reference's default backend
// (domain/math/fastmath/complex_linear_algebra.clj:174-1470): eigenvalues ascending, singular values descending,
// A = P L U, A = Q R, A = L L^H, principal branches for log / sqrt.  Eigenvector phases are unpinned by the reference
// tests; results satisfy A v = lambda v to 1e-10.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "la_host.h"

namespace qcb {
namespace la {

typedef std::complex<double> cd;

struct Mat {
  size_t r = 0, c = 0;
  std::vector<cd> a;
  Mat() {}
  Mat(size_t rr, size_t cc) : r(rr), c(cc), a(rr * cc, cd(0, 0)) {}
  cd& operator()(size_t i, size_t j) { return a[i * c + j]; }
  const cd& operator()(size_t i, size_t j) const { return a[i * c + j]; }
};

static Mat from_raw(const double* p, size_t r, size_t c) {
  Mat m(r, c);
  for (size_t i = 0; i < r * c; ++i) m.a[i] = cd(p[2 * i], p[2 * i + 1]);
  return m;
}
static void to_raw(const Mat& m, double* p) {
  for (size_t i = 0; i < m.r * m.c; ++i) { p[2 * i] = m.a[i].real(); p[2 * i + 1] = m.a[i].imag(); }
}
static Mat eye(size_t n) { Mat m(n, n); for (size_t i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
static Mat mul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (size_t i = 0; i < A.r; ++i)
    for (size_t k = 0; k < A.c; ++k) {
      const cd aik = A(i, k);
      if (aik == cd(0, 0)) continue;
      for (size_t j = 0; j < B.c; ++j) C(i, j) += aik * B(k, j);
    }
  return C;
}
static Mat adjoint(const Mat& A) {
  Mat T(A.c, A.r);
  for (size_t i = 0; i < A.r; ++i) for (size_t j = 0; j < A.c; ++j) T(j, i) = std::conj(A(i, j));
  return T;
}
static Mat add(const Mat& A, const Mat& B, cd beta = 1.0) {
  Mat C = A;
  for (size_t i = 0; i < C.a.size(); ++i) C.a[i] += beta * B.a[i];
  return C;
}
static double norm1(const Mat& A) {       // max column sum
  double best = 0;
  for (size_t j = 0; j < A.c; ++j) { double s = 0; for (size_t i = 0; i < A.r; ++i) s += std::abs(A(i, j)); best = std::max(best, s); }
  return best;
}
static double fro(const Mat& A) { double s = 0; for (const cd& z : A.a) s += std::norm(z); return std::sqrt(s); }

// ------------------------------------------------------------------ LU (partial pivoting): P A = L U, perm[i] = source row
static bool lu_factor(Mat& A, std::vector<size_t>& perm) {
  const size_t n = A.r;
  perm.resize(n);
  std::iota(perm.begin(), perm.end(), 0);
  bool singular = false;
  for (size_t k = 0; k < n; ++k) {
    size_t piv = k; double best = std::abs(A(k, k));
    for (size_t i = k + 1; i < n; ++i) if (std::abs(A(i, k)) > best) { best = std::abs(A(i, k)); piv = i; }
    if (best < 1e-300) { singular = true; continue; }
    if (piv != k) { for (size_t j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j)); std::swap(perm[k], perm[piv]); }
    for (size_t i = k + 1; i < n; ++i) {
      A(i, k) /= A(k, k);
      const cd f = A(i, k);
      for (size_t j = k + 1; j < n; ++j) A(i, j) -= f * A(k, j);
    }
  }
  return !singular;
}
static void lu_solve(const Mat& LU, const std::vector<size_t>& perm, const Mat& B, Mat& X) {
  const size_t n = LU.r, m = B.c;
  X = Mat(n, m);
  for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < m; ++j) X(i, j) = B(perm[i], j);
  for (size_t i = 0; i < n; ++i) for (size_t k = 0; k < i; ++k) for (size_t j = 0; j < m; ++j) X(i, j) -= LU(i, k) * X(k, j);
  for (size_t ii = n; ii-- > 0;) {
    for (size_t k = ii + 1; k < n; ++k) for (size_t j = 0; j < m; ++j) X(ii, j) -= LU(ii, k) * X(k, j);
    for (size_t j = 0; j < m; ++j) X(ii, j) /= LU(ii, ii);
  }
}

// ------------------------------------------------------------------ Hermitian eigenproblem: cyclic complex Jacobi
// H (overwritten) -> diagonal, V columns = eigenvectors.  Each rotation annihilates H(p,q).
static void jacobi_hermitian(Mat& H, Mat& V) {
  const size_t n = H.r;
  V = eye(n);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0;
    for (size_t p = 0; p < n; ++p) for (size_t q = p + 1; q < n; ++q) off += std::norm(H(p, q));
    if (off < 1e-300 || std::sqrt(off) < 1e-15 * std::max(1.0, fro(H))) break;
    for (size_t p = 0; p < n; ++p)
      for (size_t q = p + 1; q < n; ++q) {
        const cd c = H(p, q);
        const double ac = std::abs(c);
        if (ac < 1e-300) continue;
        const double a = H(p, p).real(), b = H(q, q).real();
        const cd ph = c / ac;                                  // e^{i phi}
        const double tau = (b - a) / (2.0 * ac);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1.0 + tau * tau));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = t * cs;
        // G = [[cs, sn*ph],[-sn*conj(ph), cs]] acting on columns (p,q): H <- G^H H G
        for (size_t k = 0; k < n; ++k) {                       // columns
          const cd hkp = H(k, p), hkq = H(k, q);
          H(k, p) = cs * hkp - sn * std::conj(ph) * hkq;
          H(k, q) = sn * ph * hkp + cs * hkq;
          const cd vkp = V(k, p), vkq = V(k, q);
          V(k, p) = cs * vkp - sn * std::conj(ph) * vkq;
          V(k, q) = sn * ph * vkp + cs * vkq;
        }
        for (size_t k = 0; k < n; ++k) {                       // rows
          const cd hpk = H(p, k), hqk = H(q, k);
          H(p, k) = cs * hpk - sn * ph * hqk;
          H(q, k) = sn * std::conj(ph) * hpk + cs * hqk;
        }
        H(p, q) = 0; H(q, p) = 0;
        H(p, p) = cd(H(p, p).real(), 0); H(q, q) = cd(H(q, q).real(), 0);
      }
  }
}

// ------------------------------------------------------------------ complex Schur form A = Q T Q^H (Hessenberg + shifted QR)
static bool schur(const Mat& A0, Mat& T, Mat& Q) {
  const size_t n = A0.r;
  T = A0; Q = eye(n);
  if (n <= 1) return true;
  // Householder reduction to upper Hessenberg form
  for (size_t k = 0; k + 2 < n; ++k) {
    double alpha = 0;
    for (size_t i = k + 1; i < n; ++i) alpha += std::norm(T(i, k));
    alpha = std::sqrt(alpha);
    if (alpha < 1e-300) continue;
    std::vector<cd> v(n, cd(0, 0));
    const cd x0 = T(k + 1, k);
    const cd phase = std::abs(x0) > 0 ? x0 / std::abs(x0) : cd(1, 0);
    for (size_t i = k + 1; i < n; ++i) v[i] = T(i, k);
    v[k + 1] += phase * alpha;
    double vn = 0; for (size_t i = k + 1; i < n; ++i) vn += std::norm(v[i]);
    if (vn < 1e-300) continue;
    for (size_t j = 0; j < n; ++j) {     // T <- (I - 2 v v^H / vn) T
      cd s = 0; for (size_t i = k + 1; i < n; ++i) s += std::conj(v[i]) * T(i, j);
      s *= 2.0 / vn;
      for (size_t i = k + 1; i < n; ++i) T(i, j) -= v[i] * s;
    }
    for (size_t i = 0; i < n; ++i) {     // T <- T (I - 2 v v^H / vn),  Q likewise
      cd s = 0; for (size_t j = k + 1; j < n; ++j) s += T(i, j) * v[j];
      s *= 2.0 / vn;
      for (size_t j = k + 1; j < n; ++j) T(i, j) -= s * std::conj(v[j]);
      cd sq = 0; for (size_t j = k + 1; j < n; ++j) sq += Q(i, j) * v[j];
      sq *= 2.0 / vn;
      for (size_t j = k + 1; j < n; ++j) Q(i, j) -= sq * std::conj(v[j]);
    }
  }
  // single-shift QR iterations with deflation (Givens rotations on the Hessenberg matrix)
  size_t hi = n - 1;
  int iter = 0;
  while (hi > 0) {
    size_t l = hi;
    while (l > 0) {
      const double s = std::abs(T(l - 1, l - 1)) + std::abs(T(l, l));
      if (std::abs(T(l, l - 1)) <= 1e-16 * (s > 0 ? s : 1.0)) { T(l, l - 1) = 0; break; }
      --l;
    }
    if (l == hi) { --hi; iter = 0; continue; }
    if (++iter > 500) return false;
    // Wilkinson shift: eigenvalue of the trailing 2x2 closer to T(hi,hi)
    const cd a = T(hi - 1, hi - 1), b = T(hi - 1, hi), c = T(hi, hi - 1), d = T(hi, hi);
    const cd tr = a + d, det = a * d - b * c;
    const cd disc = std::sqrt(tr * tr - 4.0 * det);
    cd mu1 = (tr + disc) / 2.0, mu2 = (tr - disc) / 2.0;
    cd mu = std::abs(mu1 - d) < std::abs(mu2 - d) ? mu1 : mu2;
    if (iter % 11 == 10) mu += cd(std::abs(T(hi, hi - 1)), 0);          // exceptional shift
    for (size_t i = l; i <= hi; ++i) T(i, i) -= mu;
    std::vector<cd> cs(hi - l), sn(hi - l);
    for (size_t k = l; k < hi; ++k) {                                    // QR by Givens: T = R
      const cd x = T(k, k), y = T(k + 1, k);
      const double r = std::sqrt(std::norm(x) + std::norm(y));
      cd c1 = 1.0, s1 = 0.0;
      if (r > 1e-300) { c1 = x / r; s1 = y / r; }
      cs[k - l] = c1; sn[k - l] = s1;
      for (size_t j = k; j < n; ++j) {
        const cd t1 = T(k, j), t2 = T(k + 1, j);
        T(k, j) = std::conj(c1) * t1 + std::conj(s1) * t2;
        T(k + 1, j) = -s1 * t1 + c1 * t2;
      }
    }
    for (size_t k = l; k < hi; ++k) {                                    // T = R G^H ..., Q accumulates
      const cd c1 = cs[k - l], s1 = sn[k - l];
      const size_t top = std::min(hi, k + 2);
      for (size_t i = 0; i <= top; ++i) {
        const cd t1 = T(i, k), t2 = T(i, k + 1);
        T(i, k) = t1 * c1 + t2 * s1;
        T(i, k + 1) = -t1 * std::conj(s1) + t2 * std::conj(c1);
      }
      for (size_t i = 0; i < n; ++i) {
        const cd q1 = Q(i, k), q2 = Q(i, k + 1);
        Q(i, k) = q1 * c1 + q2 * s1;
        Q(i, k + 1) = -q1 * std::conj(s1) + q2 * std::conj(c1);
      }
    }
    for (size_t i = l; i <= hi; ++i) T(i, i) += mu;
  }
  for (size_t i = 1; i < n; ++i) for (size_t j = 0; j < i; ++j) T(i, j) = 0;
  return true;
}

// ------------------------------------------------------------------ one-sided Jacobi SVD of an m x n matrix, m >= n
// On exit: A's columns = U_k * sigma_k, V = right singular vectors (n x n)
static void jacobi_svd_tall(Mat& A, Mat& V) {
  const size_t m = A.r, n = A.c;
  V = eye(n);
  for (int sweep = 0; sweep < 100; ++sweep) {
    bool rotated = false;
    for (size_t p = 0; p < n; ++p)
      for (size_t q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0; cd gamma = 0;
        for (size_t i = 0; i < m; ++i) { alpha += std::norm(A(i, p)); beta += std::norm(A(i, q)); gamma += std::conj(A(i, p)) * A(i, q); }
        const double ag = std::abs(gamma);
        if (ag <= 1e-15 * std::sqrt(alpha * beta) || ag < 1e-300) continue;
        rotated = true;
        const cd ph = gamma / ag;
        const double tau = (beta - alpha) / (2.0 * ag);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1.0 + tau * tau));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = t * cs;
        for (size_t i = 0; i < m; ++i) {
          const cd ap = A(i, p), aq = A(i, q);
          A(i, p) = cs * ap - sn * std::conj(ph) * aq;
          A(i, q) = sn * ph * ap + cs * aq;
        }
        for (size_t i = 0; i < n; ++i) {
          const cd vp = V(i, p), vq = V(i, q);
          V(i, p) = cs * vp - sn * std::conj(ph) * vq;
          V(i, q) = sn * ph * vp + cs * vq;
        }
      }
    if (!rotated) break;
  }
}

// complete the first k orthonormal columns of U (m x m) to a unitary matrix (modified Gram-Schmidt on unit vectors)
static void complete_unitary(Mat& U, size_t k) {
  const size_t m = U.r;
  for (size_t col = k; col < m; ++col) {
    for (size_t trial = 0; trial < m; ++trial) {
      std::vector<cd> v(m, cd(0, 0));
      v[(col + trial) % m] = 1.0;
      for (int pass = 0; pass < 2; ++pass)
        for (size_t j = 0; j < col; ++j) {
          cd d = 0; for (size_t i = 0; i < m; ++i) d += std::conj(U(i, j)) * v[i];
          for (size_t i = 0; i < m; ++i) v[i] -= d * U(i, j);
        }
      double nn = 0; for (const cd& z : v) nn += std::norm(z);
      if (nn > 1e-6) { nn = std::sqrt(nn); for (size_t i = 0; i < m; ++i) U(i, col) = v[i] / nn; break; }
    }
  }
}

static void svd_full(const Mat& A, Mat& U, std::vector<double>& S, Mat& Vh) {
  const bool wide = A.r < A.c;
  Mat W = wide ? adjoint(A) : A;                 // tall
  const size_t m = W.r, n = W.c;
  Mat V;
  jacobi_svd_tall(W, V);
  std::vector<double> sig(n);
  for (size_t j = 0; j < n; ++j) { double s = 0; for (size_t i = 0; i < m; ++i) s += std::norm(W(i, j)); sig[j] = std::sqrt(s); }
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return sig[a] > sig[b]; });
  Mat Ut(m, m), Vs(n, n);
  S.assign(n, 0.0);
  size_t rank = 0;
  const double tol = (n ? sig[order[0]] : 0.0) * 1e-14 * (double)std::max(m, n);
  for (size_t k = 0; k < n; ++k) {
    const size_t j = order[k];
    S[k] = sig[j];
    for (size_t i = 0; i < n; ++i) Vs(i, k) = V(i, j);
    if (sig[j] > tol && sig[j] > 0) { for (size_t i = 0; i < m; ++i) Ut(i, k) = W(i, j) / sig[j]; rank = k + 1; }
  }
  // columns of Ut beyond `rank` (zero singular values) and beyond n: any orthonormal completion
  if (rank < n) {                                  // keep the already valid columns contiguous: they are (sorted descending)
    Mat tmp(m, m);
    for (size_t k = 0; k < rank; ++k) for (size_t i = 0; i < m; ++i) tmp(i, k) = Ut(i, k);
    Ut = tmp;
  }
  complete_unitary(Ut, rank);
  if (!wide) { U = Ut; Vh = adjoint(Vs); }
  else { U = Vs; Vh = adjoint(Ut); }               // A = (W)^H = Vs S Ut^H
}

// ------------------------------------------------------------------ functions of matrices
static Mat expm(const Mat& A) {
  const size_t n = A.r;
  int s = 0;
  double nrm = norm1(A);
  while (nrm > 0.25) { nrm *= 0.5; ++s; }
  Mat X = A;
  const double sc = std::ldexp(1.0, -s);
  for (cd& z : X.a) z *= sc;
  Mat R = eye(n), term = eye(n);
  for (int k = 1; k <= 24; ++k) {
    term = mul(term, X);
    for (cd& z : term.a) z /= (double)k;
    R = add(R, term);
    if (fro(term) < 1e-18 * std::max(1.0, fro(R))) break;
  }
  for (int i = 0; i < s; ++i) R = mul(R, R);
  return R;
}

// principal square root of an upper-triangular matrix (Bjorck-Hammarling)
static bool sqrt_triangular(const Mat& T, Mat& R) {
  const size_t n = T.r;
  R = Mat(n, n);
  for (size_t i = 0; i < n; ++i) R(i, i) = std::sqrt(T(i, i));
  for (size_t d = 1; d < n; ++d)
    for (size_t i = 0; i + d < n; ++i) {
      const size_t j = i + d;
      cd s = T(i, j);
      for (size_t k = i + 1; k < j; ++k) s -= R(i, k) * R(k, j);
      const cd den = R(i, i) + R(j, j);
      if (std::abs(den) < 1e-300) { if (std::abs(s) > 1e-12) return false; R(i, j) = 0; }
      else R(i, j) = s / den;
    }
  return true;
}

static bool sqrtm(const Mat& A, Mat& out) {
  Mat T, Q, R;
  if (!schur(A, T, Q)) return false;
  if (!sqrt_triangular(T, R)) return false;
  out = mul(mul(Q, R), adjoint(Q));
  return true;
}

static bool logm(const Mat& A, Mat& out) {
  const size_t n = A.r;
  Mat T, Q;
  if (!schur(A, T, Q)) return false;
  for (size_t i = 0; i < n; ++i) if (std::abs(T(i, i)) < 1e-300) return false;       // singular: no logarithm
  int k = 0;
  Mat I = eye(n);
  while (fro(add(T, I, -1.0)) > 0.25 && k < 60) {       // inverse scaling and squaring on the triangular factor
    Mat R;
    if (!sqrt_triangular(T, R)) return false;
    T = R; ++k;
  }
  Mat X = add(T, I, -1.0), term = X, L = X;            // log(I + X) = X - X^2/2 + X^3/3 - ...
  for (int j = 2; j <= 60; ++j) {
    term = mul(term, X);
    Mat t = term;
    const double c = ((j & 1) ? 1.0 : -1.0) / (double)j;
    for (cd& z : t.a) z *= c;
    L = add(L, t);
    if (fro(t) < 1e-18 * std::max(1e-300, fro(L))) break;
  }
  const double sc = std::ldexp(1.0, k);
  for (cd& z : L.a) z *= sc;
  out = mul(mul(Q, L), adjoint(Q));
  return true;
}

}  // namespace la

using namespace la;

// ------------------------------------------------------------------ entry points used by sim.cu's extern "C" wrappers
int la_hadamard(const double* A, const double* B, uint64_t n, double* C) {
  for (uint64_t i = 0; i < n; ++i) {
    const cd z = cd(A[2 * i], A[2 * i + 1]) * cd(B[2 * i], B[2 * i + 1]);
    C[2 * i] = z.real(); C[2 * i + 1] = z.imag();
  }
  return 0;
}
int la_transpose(const double* A, uint64_t rows, uint64_t cols, int conjugate, double* out) {
  for (uint64_t i = 0; i < rows; ++i)
    for (uint64_t j = 0; j < cols; ++j) {
      out[2 * (j * rows + i)] = A[2 * (i * cols + j)];
      out[2 * (j * rows + i) + 1] = conjugate ? -A[2 * (i * cols + j) + 1] : A[2 * (i * cols + j) + 1];
    }
  return 0;
}
int la_solve(const double* A, const double* B, uint64_t n, uint64_t nrhs, double* X, std::string& err) {
  Mat LU = from_raw(A, n, n), Bm = from_raw(B, n, nrhs), Xm;
  std::vector<size_t> perm;
  if (!lu_factor(LU, perm)) { err = "Matrix is singular"; return -1; }
  lu_solve(LU, perm, Bm, Xm);
  to_raw(Xm, X);
  return 0;
}
int la_inverse(const double* A, uint64_t n, double* out, std::string& err) {
  Mat LU = from_raw(A, n, n), Xm;
  std::vector<size_t> perm;
  if (!lu_factor(LU, perm)) { err = "Matrix is singular"; return -1; }
  lu_solve(LU, perm, eye(n), Xm);
  to_raw(Xm, out);
  return 0;
}
int la_is_hermitian(const double* A, uint64_t n, double eps) {
  Mat M = from_raw(A, n, n);
  for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) if (!(std::abs(M(i, j) - std::conj(M(j, i))) < eps)) return 0;
  return 1;
}
int la_is_diagonal(const double* A, uint64_t n, double eps) {
  Mat M = from_raw(A, n, n);
  for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) if (i != j && !(std::abs(M(i, j)) < eps)) return 0;
  return 1;
}
int la_is_unitary(const double* A, uint64_t n, double eps) {
  Mat M = from_raw(A, n, n), P = mul(adjoint(M), M);
  for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) if (!(std::abs(P(i, j) - (i == j ? cd(1, 0) : cd(0, 0))) < eps)) return 0;
  return 1;
}
int la_eigh(const double* A, uint64_t n, double* evals, double* evecs) {
  Mat H = from_raw(A, n, n), V;
  for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) { const cd m = 0.5 * (H(i, j) + std::conj(H(j, i))); H(i, j) = m; H(j, i) = std::conj(m); }
  jacobi_hermitian(H, V);
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return H(a, a).real() < H(b, b).real(); });
  for (size_t k = 0; k < n; ++k) {
    evals[k] = H(order[k], order[k]).real();
    for (size_t i = 0; i < n; ++i) { evecs[2 * (k * n + i)] = V(i, order[k]).real(); evecs[2 * (k * n + i) + 1] = V(i, order[k]).imag(); }
  }
  return 0;
}
int la_is_psd(const double* A, uint64_t n, double eps, std::string& err) {
  if (!la_is_hermitian(A, n, eps)) { err = "Matrix must be Hermitian for positive-semidefinite check"; return -1; }
  std::vector<double> ev(n), vec(2 * n * n);
  la_eigh(A, n, ev.data(), vec.data());
  for (double e : ev) if (!(e >= -eps)) return 0;
  return 1;
}
int la_eig(const double* A, uint64_t n, double* evals, double* evecs, std::string& err) {
  Mat T, Q;
  if (!schur(from_raw(A, n, n), T, Q)) { err = "QR iteration did not converge"; return -1; }
  // eigenvectors of the triangular factor by back substitution, then rotate by Q
  Mat Y(n, n);
  const double small = 1e-14 * std::max(1.0, fro(T));
  for (size_t k = 0; k < n; ++k) {
    Y(k, k) = 1.0;
    for (size_t ii = k; ii-- > 0;) {
      cd s = 0;
      for (size_t j = ii + 1; j <= k; ++j) s += T(ii, j) * Y(j, k);
      cd den = T(ii, ii) - T(k, k);
      if (std::abs(den) < small) den = small;          // (nearly) repeated eigenvalue: perturb, as LAPACK's trevc does
      Y(ii, k) = -s / den;
    }
  }
  Mat V = mul(Q, Y);
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    const cd x = T(a, a), y = T(b, b);
    return x.real() != y.real() ? x.real() < y.real() : x.imag() < y.imag();
  });
  for (size_t k = 0; k < n; ++k) {
    const size_t j = order[k];
    evals[2 * k] = T(j, j).real(); evals[2 * k + 1] = T(j, j).imag();
    double nn = 0; for (size_t i = 0; i < n; ++i) nn += std::norm(V(i, j));
    nn = nn > 0 ? std::sqrt(nn) : 1.0;
    for (size_t i = 0; i < n; ++i) { evecs[2 * (k * n + i)] = V(i, j).real() / nn; evecs[2 * (k * n + i) + 1] = V(i, j).imag() / nn; }
  }
  return 0;
}
int la_svd(const double* A, uint64_t m, uint64_t n, double* U, double* S, double* Vh) {
  Mat Um, Vhm; std::vector<double> s;
  svd_full(from_raw(A, m, n), Um, s, Vhm);
  to_raw(Um, U); to_raw(Vhm, Vh);
  for (size_t i = 0; i < s.size(); ++i) S[i] = s[i];
  return 0;
}
int la_lu(const double* A, uint64_t n, double* P, double* L, double* U) {
  Mat LU = from_raw(A, n, n);
  std::vector<size_t> perm;
  lu_factor(LU, perm);                              // P_row A = L U with P_row[i][perm[i]] = 1  =>  A = P L U with P = P_row^T
  Mat Pm(n, n), Lm = eye(n), Um(n, n);
  for (size_t i = 0; i < n; ++i) {
    Pm(perm[i], i) = 1.0;
    for (size_t j = 0; j < n; ++j) { if (j < i) Lm(i, j) = LU(i, j); else Um(i, j) = LU(i, j); }
  }
  to_raw(Pm, P); to_raw(Lm, L); to_raw(Um, U);
  return 0;
}
int la_qr(const double* A, uint64_t m, uint64_t n, double* Qo, double* Ro) {
  Mat R = from_raw(A, m, n), Q = eye(m);
  const size_t steps = std::min<size_t>(m > 0 ? m - 1 : 0, n);
  for (size_t k = 0; k < steps; ++k) {
    double alpha = 0; for (size_t i = k; i < m; ++i) alpha += std::norm(R(i, k));
    alpha = std::sqrt(alpha);
    if (alpha < 1e-300) continue;
    std::vector<cd> v(m, cd(0, 0));
    const cd x0 = R(k, k);
    const cd phase = std::abs(x0) > 0 ? x0 / std::abs(x0) : cd(1, 0);
    for (size_t i = k; i < m; ++i) v[i] = R(i, k);
    v[k] += phase * alpha;
    double vn = 0; for (size_t i = k; i < m; ++i) vn += std::norm(v[i]);
    if (vn < 1e-300) continue;
    for (size_t j = 0; j < n; ++j) {
      cd s = 0; for (size_t i = k; i < m; ++i) s += std::conj(v[i]) * R(i, j);
      s *= 2.0 / vn;
      for (size_t i = k; i < m; ++i) R(i, j) -= v[i] * s;
    }
    for (size_t i = 0; i < m; ++i) {
      cd s = 0; for (size_t j = k; j < m; ++j) s += Q(i, j) * v[j];
      s *= 2.0 / vn;
      for (size_t j = k; j < m; ++j) Q(i, j) -= s * std::conj(v[j]);
    }
  }
  for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < n && j < i; ++j) R(i, j) = 0;
  to_raw(Q, Qo); to_raw(R, Ro);
  return 0;
}
int la_cholesky(const double* A, uint64_t n, double* Lo, std::string& err) {
  Mat M = from_raw(A, n, n), L(n, n);
  for (size_t j = 0; j < n; ++j) {
    double d = M(j, j).real();
    for (size_t k = 0; k < j; ++k) d -= std::norm(L(j, k));
    if (d < -1e-12 * std::max(1.0, std::abs(M(j, j)))) { err = "Matrix is not positive semidefinite"; return -1; }
    const double ljj = std::sqrt(std::max(d, 0.0));
    L(j, j) = ljj;
    for (size_t i = j + 1; i < n; ++i) {
      cd s = M(i, j);
      for (size_t k = 0; k < j; ++k) s -= L(i, k) * std::conj(L(j, k));
      L(i, j) = ljj > 1e-300 ? s / ljj : cd(0, 0);
    }
  }
  to_raw(L, Lo);
  return 0;
}
int la_expm(const double* A, uint64_t n, double* out) { to_raw(expm(from_raw(A, n, n)), out); return 0; }
int la_logm(const double* A, uint64_t n, double* out, std::string& err) {
  Mat R;
  if (!logm(from_raw(A, n, n), R)) { err = "matrix logarithm does not exist (singular matrix) or did not converge"; return -1; }
  to_raw(R, out);
  return 0;
}
int la_sqrtm(const double* A, uint64_t n, double* out, std::string& err) {
  Mat R;
  if (!sqrtm(from_raw(A, n, n), R)) { err = "matrix square root does not exist or did not converge"; return -1; }
  to_raw(R, out);
  return 0;
}
int la_singular_values(const double* A, uint64_t m, uint64_t n, std::vector<double>& s) {
  Mat U, Vh;
  svd_full(from_raw(A, m, n), U, s, Vh);
  return 0;
}

}  // namespace qcb
