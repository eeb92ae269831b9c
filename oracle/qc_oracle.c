/*
 * qc_oracle.c — plain-C CPU restatement of the reference's state-vector hot path.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/_build/libqcoracle.so and loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product library
 * (libqcb200.so) never links or loads it.
 *
 * It mirrors oracle/qc_oracle.py (which is pinned against the reference's golden vectors) function by
 * function at C speed, so that 20-30 qubit circuits can be cross-checked and timed on the host cores.
 * Citations are relative to /root/reference/src/org/soulspace/qclojure/.  Semantics are the
 * reference's (strict parity): qubit 0 = MSB (domain/state.clj:114-162); controlled gates apply the
 * transposed 2x2 (domain/gate.clj:473-483); SWAP/iSWAP operands count from the LSB (gate.clj:768-778).
 *
 * Two modes:
 *   orc_apply_ops            in-place pairwise update, OpenMP over all host cores ("cpu-stride", the fair
 *                            O(2^n) CPU comparison of SURVEY §8d)
 *   orc_apply_1q_dense_kron  the reference's own algorithm: expand the gate to a dense 2^n x 2^n matrix by
 *                            Kronecker products and do a dense mat-vec (gate.clj:346-353, 384-395;
 *                            math/fastmath/complex_linear_algebra.clj:408-431, 463-484), single thread
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/qcb200.h"

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) { cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }

/* insert a zero bit at position p into k */
static inline uint64_t ins0(uint64_t k, int p) { return ((k >> p) << (p + 1)) | (k & ((1ULL << p) - 1)); }

/* gate.clj:384-395 pairwise form: new0 = U00 a0 + U01 a1 ; new1 = U10 a0 + U11 a1, on index bit `bit` */
static void apply_1q_bit(cplx* s, int n, int bit, const cplx U[4]) {
  const uint64_t half = 1ULL << (n - 1), st = 1ULL << bit;
#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < half; ++k) {
    uint64_t i0 = ins0(k, bit), i1 = i0 | st;
    cplx a0 = s[i0], a1 = s[i1];
    s[i0] = cadd(cmul(U[0], a0), cmul(U[1], a1));
    s[i1] = cadd(cmul(U[2], a0), cmul(U[3], a1));
  }
}

/* gate.clj:451-486: where control bit = 1: new0 = U00 a0 + U10 a1 ; new1 = U01 a0 + U11 a1 (U^T) */
static void apply_ctrl_bit(cplx* s, int n, int cbit, int tbit, const cplx U[4], int transpose) {
  const uint64_t quarter = 1ULL << (n - 2);
  int lo = cbit < tbit ? cbit : tbit, hi = cbit < tbit ? tbit : cbit;
  cplx u01 = transpose ? U[2] : U[1], u10 = transpose ? U[1] : U[2];
#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < quarter; ++k) {
    uint64_t b = ins0(ins0(k, lo), hi) | (1ULL << cbit);
    uint64_t i0 = b, i1 = b | (1ULL << tbit);
    cplx a0 = s[i0], a1 = s[i1];
    s[i0] = cadd(cmul(a0, U[0]), cmul(a1, u01));
    s[i1] = cadd(cmul(a0, u10), cmul(a1, U[3]));
  }
}

/* swap / iswap on raw index bits b1,b2 with optional control bit (fredkin) */
static void apply_swap_bits(cplx* s, int n, int b1, int b2, int cbit, cplx phase, int has_phase) {
  const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; ++i) {
    if (cbit >= 0 && !((i >> cbit) & 1)) continue;
    if (((i >> b1) & 1) == 0 && ((i >> b2) & 1) == 1) {
      uint64_t j = (i | (1ULL << b1)) & ~(1ULL << b2);
      cplx a = s[i], b = s[j];
      if (has_phase) { a = cmul(phase, a); b = cmul(phase, b); }
      s[i] = b; s[j] = a;
    }
  }
}

static void apply_phase_mask(cplx* s, int n, uint64_t mask, cplx ph) {
  const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; ++i)
    if ((i & mask) == mask) s[i] = cmul(s[i], ph);
}

static void apply_phase_pop1(cplx* s, int n, uint64_t mask, cplx ph) {
  const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; ++i)
    if (__builtin_popcountll(i & mask) == 1) s[i] = cmul(s[i], ph);
}

static void mat_rx(double t, cplx U[4]) { double c = cos(t / 2), s = sin(t / 2); U[0] = (cplx){c, 0}; U[1] = (cplx){0, -s}; U[2] = (cplx){0, -s}; U[3] = (cplx){c, 0}; }
static void mat_ry(double t, cplx U[4]) { double c = cos(t / 2), s = sin(t / 2); U[0] = (cplx){c, 0}; U[1] = (cplx){-s, 0}; U[2] = (cplx){s, 0}; U[3] = (cplx){c, 0}; }
static void mat_rz(double t, cplx U[4]) { U[0] = (cplx){cos(t / -2), sin(t / -2)}; U[1] = (cplx){0, 0}; U[2] = (cplx){0, 0}; U[3] = (cplx){cos(t / 2), sin(t / 2)}; }
static void mat_phase(double p, cplx U[4]) { U[0] = (cplx){1, 0}; U[1] = (cplx){0, 0}; U[2] = (cplx){0, 0}; U[3] = (cplx){cos(p), sin(p)}; }

static const cplx MX[4] = {{0,0},{1,0},{1,0},{0,0}};
static const cplx MY[4] = {{0,0},{0,-1},{0,1},{0,0}};
static const cplx MZ[4] = {{1,0},{0,0},{0,0},{-1,0}};
static const cplx MS[4] = {{1,0},{0,0},{0,0},{0,1}};
static const cplx MSD[4] = {{1,0},{0,0},{0,0},{0,-1}};

/* circuit.clj:952-1072 dispatch on the qcb_op encoding.  Returns 0, or QCB_ERR_UNSUPPORTED. */
int orc_apply_ops(double* state, int n, const qcb_op* ops, uint64_t n_ops) {
  cplx* s = (cplx*)state;
  for (uint64_t k = 0; k < n_ops; ++k) {
    const qcb_op* op = &ops[k];
    cplx U[4];
    const double r2 = 1.0 / sqrt(2.0);
    int q0 = op->q[0], q1 = op->q[1], q2 = op->q[2];
#define BIT(q) (n - 1 - (q))
    switch (op->kind) {
      case QCB_OP_X: apply_1q_bit(s, n, BIT(q0), MX); break;
      case QCB_OP_Y: apply_1q_bit(s, n, BIT(q0), MY); break;
      case QCB_OP_Z: apply_1q_bit(s, n, BIT(q0), MZ); break;
      case QCB_OP_H: U[0] = (cplx){r2,0}; U[1] = (cplx){r2,0}; U[2] = (cplx){r2,0}; U[3] = (cplx){-r2,0}; apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_S: apply_1q_bit(s, n, BIT(q0), MS); break;
      case QCB_OP_SDG: apply_1q_bit(s, n, BIT(q0), MSD); break;
      case QCB_OP_T: mat_phase(M_PI / 4, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_TDG: mat_phase(M_PI / -4, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_RX: mat_rx(op->angle, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_RY: mat_ry(op->angle, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_RZ: mat_rz(op->angle, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_PHASE: mat_phase(op->angle, U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_CNOT: apply_ctrl_bit(s, n, BIT(q0), BIT(q1), MX, 1); break;
      case QCB_OP_CZ: case QCB_OP_RYDBERG_CZ: apply_ctrl_bit(s, n, BIT(q0), BIT(q1), MZ, 1); break;
      case QCB_OP_CRX: mat_rx(op->angle, U); apply_ctrl_bit(s, n, BIT(q0), BIT(q1), U, 1); break;
      case QCB_OP_CRY: mat_ry(op->angle, U); apply_ctrl_bit(s, n, BIT(q0), BIT(q1), U, 1); break;
      case QCB_OP_CRZ: mat_rz(op->angle, U); apply_ctrl_bit(s, n, BIT(q0), BIT(q1), U, 1); break;
      case QCB_OP_SWAP: apply_swap_bits(s, n, q0, q1, -1, (cplx){1, 0}, 0); break;       /* LSB positions */
      case QCB_OP_ISWAP: apply_swap_bits(s, n, q0, q1, -1, (cplx){0, 1}, 1); break;
      case QCB_OP_TOFFOLI: {  /* gate.clj:869-900 */
        uint64_t cm = (1ULL << BIT(q0)) | (1ULL << BIT(q1)), tb = 1ULL << BIT(q2), N = 1ULL << n;
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < N; ++i)
          if ((i & cm) == cm && !(i & tb)) { cplx a = s[i]; s[i] = s[i | tb]; s[i | tb] = a; }
        break; }
      case QCB_OP_FREDKIN: apply_swap_bits(s, n, BIT(q1), BIT(q2), BIT(q0), (cplx){1, 0}, 0); break;
      case QCB_OP_RYDBERG_CPHASE: apply_phase_mask(s, n, (1ULL << BIT(q0)) | (1ULL << BIT(q1)), (cplx){cos(op->angle), sin(op->angle)}); break;
      case QCB_OP_RYDBERG_BLOCKADE: {
        uint64_t m = 0; for (int q = 0; q < n; ++q) if ((op->mask >> q) & 1) m |= 1ULL << BIT(q);
        apply_phase_pop1(s, n, m, (cplx){cos(op->angle), sin(op->angle)}); break; }
      case QCB_OP_GLOBAL_H: U[0] = (cplx){r2,0}; U[1] = (cplx){r2,0}; U[2] = (cplx){r2,0}; U[3] = (cplx){-r2,0};
        for (int q = 0; q < n; ++q) { apply_1q_bit(s, n, BIT(q), U); } break;
      case QCB_OP_GLOBAL_X: case QCB_OP_GLOBAL_RX: mat_rx(op->kind == QCB_OP_GLOBAL_X ? M_PI : op->angle, U);
        for (int q = 0; q < n; ++q) { apply_1q_bit(s, n, BIT(q), U); } break;
      case QCB_OP_GLOBAL_Y: case QCB_OP_GLOBAL_RY: mat_ry(op->kind == QCB_OP_GLOBAL_Y ? M_PI : op->angle, U);
        for (int q = 0; q < n; ++q) { apply_1q_bit(s, n, BIT(q), U); } break;
      case QCB_OP_GLOBAL_Z: case QCB_OP_GLOBAL_RZ: mat_rz(op->kind == QCB_OP_GLOBAL_Z ? M_PI : op->angle, U);
        for (int q = 0; q < n; ++q) { apply_1q_bit(s, n, BIT(q), U); } break;
      case QCB_OP_U1Q: memcpy(U, op->mat, sizeof U); apply_1q_bit(s, n, BIT(q0), U); break;
      case QCB_OP_CU1Q: memcpy(U, op->mat, sizeof U); apply_ctrl_bit(s, n, BIT(q0), BIT(q1), U, 0); break;
      default: return QCB_ERR_UNSUPPORTED;   /* :i, :cy -> "Unknown gate type", circuit.clj:1072 */
    }
#undef BIT
  }
  return 0;
}

/* state.clj:676-682 — p_i = (hypot(re,im))^2 */
void orc_probabilities(const double* state, int n, double* out) {
  const cplx* s = (const cplx*)state; const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; ++i) { double m = hypot(s[i].re, s[i].im); out[i] = m * m; }
}

/* state.clj:894-913 — sequential running sum, outcome = #{i : cum_i < total*u}, clamped.  cum: scratch[2^n]. */
int orc_sample(const double* state, int n, const double* uniforms, uint64_t shots, uint64_t* outcomes, double* cum) {
  const uint64_t N = 1ULL << n;
  orc_probabilities(state, n, cum);
  double acc = 0; for (uint64_t i = 0; i < N; ++i) { acc += cum[i]; cum[i] = acc; }
  double total = acc;
  if (fabs(total - 1.0) > 1e-8) return QCB_ERR_STATE;
  for (uint64_t k = 0; k < shots; ++k) {
    double r = total * uniforms[k];
    uint64_t lo = 0, hi = N;                 /* first i with cum_i >= r */
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (cum[mid] < r) lo = mid + 1; else hi = mid; }
    outcomes[k] = lo < N - 1 ? lo : N - 1;
  }
  return 0;
}

double orc_norm2(const double* state, int n) {
  const cplx* s = (const cplx*)state; const uint64_t N = 1ULL << n; double acc = 0;
#pragma omp parallel for reduction(+:acc) schedule(static)
  for (uint64_t i = 0; i < N; ++i) acc += s[i].re * s[i].re + s[i].im * s[i].im;
  return sqrt(acc);
}

/* observables.clj:216-251 — Re <psi|P|psi>, string char k <-> qubit k (MSB first) */
double orc_expect_pauli(const double* state, int n, const char* pauli) {
  const cplx* s = (const cplx*)state; const uint64_t N = 1ULL << n;
  uint64_t xm = 0, zm = 0; int ny = 0;
  for (int q = 0; q < n; ++q) {
    char c = pauli[q]; uint64_t b = 1ULL << (n - 1 - q);
    if (c == 'X') xm |= b; else if (c == 'Z') zm |= b; else if (c == 'Y') { xm |= b; zm |= b; ++ny; }
  }
  /* (P psi)_i = i^ny * (-1)^popcount(j & zm) * psi_j, j = i ^ xm  (Y = i X Z: Y|b> = i(-1)^b |b^1>) */
  double re = 0;
  static const double PH[4][2] = {{1,0},{0,1},{-1,0},{0,-1}};
  double pr = PH[ny & 3][0], pi = PH[ny & 3][1];
#pragma omp parallel for reduction(+:re) schedule(static)
  for (uint64_t i = 0; i < N; ++i) {
    uint64_t j = i ^ xm;
    double sg = (__builtin_popcountll(j & zm) & 1) ? -1.0 : 1.0;
    cplx v = { sg * (pr * s[j].re - pi * s[j].im), sg * (pr * s[j].im + pi * s[j].re) };
    re += s[i].re * v.re + s[i].im * v.im;     /* Re(conj(a_i) * v) */
  }
  return re;
}

/* The reference's own algorithm for a 1q gate (gate.clj:346-353 + dense mat-vec): O(4^n) time and memory.
 * Returns 0, or QCB_ERR_NOMEM when the 2^n x 2^n matrix cannot be allocated. */
int orc_apply_1q_dense_kron(double* state, int n, int target, const double mat[8]) {
  const uint64_t N = 1ULL << n;
  cplx* full = (cplx*)malloc(sizeof(cplx) * 1); if (!full) return QCB_ERR_NOMEM;
  full[0] = (cplx){1, 0}; uint64_t dim = 1;
  const cplx I2[4] = {{1,0},{0,0},{0,0},{1,0}}; const cplx* U = (const cplx*)mat;
  for (int q = 0; q < n; ++q) {                      /* reduce kronecker-product, qubit 0 leftmost */
    const cplx* B = (q == target) ? U : I2;
    uint64_t nd = dim * 2;
    cplx* nx = (cplx*)malloc(sizeof(cplx) * nd * nd); if (!nx) { free(full); return QCB_ERR_NOMEM; }
    for (uint64_t r = 0; r < dim; ++r) for (uint64_t c = 0; c < dim; ++c) {
      cplx a = full[r * dim + c];
      for (int br = 0; br < 2; ++br) for (int bc = 0; bc < 2; ++bc)
        nx[(r * 2 + br) * nd + (c * 2 + bc)] = cmul(a, B[br * 2 + bc]);
    }
    free(full); full = nx; dim = nd;
  }
  cplx* s = (cplx*)state; cplx* out = (cplx*)malloc(sizeof(cplx) * N); if (!out) { free(full); return QCB_ERR_NOMEM; }
  for (uint64_t r = 0; r < N; ++r) { cplx acc = {0, 0}; for (uint64_t c = 0; c < N; ++c) acc = cadd(acc, cmul(full[r * N + c], s[c])); out[r] = acc; }
  memcpy(s, out, sizeof(cplx) * N); free(out); free(full);
  return 0;
}

/* host threads the OpenMP loops use (torchrun exports OMP_NUM_THREADS=1 to its children: the bench's CPU legs reset it) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
