// la_host.h — host-side dense complex linear algebra behind the qcb_la_* decomposition / function entry points
// (la_host.cpp).  Raw interleaved row-major buffers; return 0 on success, -1 with `err` set otherwise.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace qcb {

int la_hadamard(const double* A, const double* B, uint64_t n, double* C);
int la_transpose(const double* A, uint64_t rows, uint64_t cols, int conjugate, double* out);
int la_solve(const double* A, const double* B, uint64_t n, uint64_t nrhs, double* X, std::string& err);
int la_inverse(const double* A, uint64_t n, double* out, std::string& err);
int la_is_hermitian(const double* A, uint64_t n, double eps);
int la_is_diagonal(const double* A, uint64_t n, double eps);
int la_is_unitary(const double* A, uint64_t n, double eps);
int la_is_psd(const double* A, uint64_t n, double eps, std::string& err);        // 1 / 0, -1 = not Hermitian
int la_eigh(const double* A, uint64_t n, double* evals, double* evecs);           // evecs: vector k at [k*n, (k+1)*n)
int la_eig(const double* A, uint64_t n, double* evals, double* evecs, std::string& err);
int la_svd(const double* A, uint64_t m, uint64_t n, double* U, double* S, double* Vh);
int la_lu(const double* A, uint64_t n, double* P, double* L, double* U);
int la_qr(const double* A, uint64_t m, uint64_t n, double* Q, double* R);
int la_cholesky(const double* A, uint64_t n, double* L, std::string& err);
int la_expm(const double* A, uint64_t n, double* out);
int la_logm(const double* A, uint64_t n, double* out, std::string& err);
int la_sqrtm(const double* A, uint64_t n, double* out, std::string& err);
int la_singular_values(const double* A, uint64_t m, uint64_t n, std::vector<double>& s);

}  // namespace qcb
