// plan.cpp — lowering of the reference gate vocabulary, the fusion scheduler and the program encoder.
// Host-only (compiled by nvcc into libqcb200.so and by g++ into the CPU test emulator).
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <system_error>
#include <thread>

namespace qcb {

static const double kPi = 3.14159265358979323846;

uint64_t Gate::target_mask() const {
  switch (kind) {
    case G_MAT1: return 1ULL << t0;
    case G_MAT2: case G_SWAPP: return (1ULL << t0) | (1ULL << t1);
    default: return 0;
  }
}
uint64_t Gate::diag_mask() const {
  switch (kind) {
    case G_MAT1: case G_MAT2: case G_SWAPP: return cmask;
    case G_DMASK: case G_DPOP1: return dmask;
    case G_DTAB1: return cmask | (1ULL << t0);
    default: return 0;
  }
}

// ------------------------------------------------------------------ matrices (domain/gate.clj:38-283)
void gate_matrix(int kind, double a, cplx U[4]) {
  const double r2 = 1.0 / std::sqrt(2.0);
  auto set = [&](cplx a00, cplx a01, cplx a10, cplx a11) { U[0] = a00; U[1] = a01; U[2] = a10; U[3] = a11; };
  double c = std::cos(a / 2), s = std::sin(a / 2);
  switch (kind) {
    case QCB_OP_X: set({0, 0}, {1, 0}, {1, 0}, {0, 0}); break;
    case QCB_OP_Y: set({0, 0}, {0, -1}, {0, 1}, {0, 0}); break;
    case QCB_OP_Z: set({1, 0}, {0, 0}, {0, 0}, {-1, 0}); break;
    case QCB_OP_H: set({r2, 0}, {r2, 0}, {r2, 0}, {-r2, 0}); break;
    case QCB_OP_S: set({1, 0}, {0, 0}, {0, 0}, {0, 1}); break;
    case QCB_OP_SDG: set({1, 0}, {0, 0}, {0, 0}, {0, -1}); break;
    case QCB_OP_T: set({1, 0}, {0, 0}, {0, 0}, {std::cos(kPi / 4), std::sin(kPi / 4)}); break;
    case QCB_OP_TDG: set({1, 0}, {0, 0}, {0, 0}, {std::cos(kPi / -4), std::sin(kPi / -4)}); break;
    case QCB_OP_PHASE: set({1, 0}, {0, 0}, {0, 0}, {std::cos(a), std::sin(a)}); break;
    case QCB_OP_RX: case QCB_OP_CRX: case QCB_OP_GLOBAL_RX: set({c, 0}, {0, -s}, {0, -s}, {c, 0}); break;
    case QCB_OP_RY: case QCB_OP_CRY: case QCB_OP_GLOBAL_RY: set({c, 0}, {-s, 0}, {s, 0}, {c, 0}); break;
    case QCB_OP_RZ: case QCB_OP_CRZ: case QCB_OP_GLOBAL_RZ:
      set({std::cos(a / -2), std::sin(a / -2)}, {0, 0}, {0, 0}, {std::cos(a / 2), std::sin(a / 2)}); break;
    default: set({1, 0}, {0, 0}, {0, 0}, {1, 0}); break;
  }
}

Config config_from(const qcb_config& c) {
  Config k;
  k.n_total = c.n_qubits;
  k.world = c.world_size > 0 ? c.world_size : 1;
  k.rank = c.rank;
  int p = 0;
  while ((1 << p) < k.world) ++p;
  k.n_local = c.n_qubits - p;
  k.tile_bits = c.tile_bits > 0 ? c.tile_bits : 12;
  k.low_bits = c.low_bits > 0 ? c.low_bits : 4;
  if (k.tile_bits > MAX_TILE_BITS) k.tile_bits = MAX_TILE_BITS;
  if (k.tile_bits > k.n_local) k.tile_bits = k.n_local;
  if (k.low_bits > k.tile_bits) k.low_bits = k.tile_bits;
  // a tile smaller than the slice must have room for the two targets of a swap / dense two-qubit gate next to its fixed low bits
  if (k.tile_bits < k.n_local) {
    if (k.tile_bits < 2) k.tile_bits = std::min(2, k.n_local);
    if (k.tile_bits < k.n_local && k.low_bits > k.tile_bits - 2) k.low_bits = k.tile_bits - 2;
  }
  k.fusion = c.fusion;
  k.strict = c.strict_parity;
  k.max_stage_cost = c.max_stage_cost;
  k.max_stage_rounds = c.max_stage_rounds;
  k.dense_mma = (c.dense_mma == 2) ? 0 : 1;
  k.mma_form = (c.dense_mma == 3) ? 1 : 0;
  if (const char* e = std::getenv("QCB_MMA_FORM")) k.mma_form = std::atoi(e) ? 1 : 0;      // experiment knob
  if (const char* e = std::getenv("QCB_DIRECT_STORE")) k.direct_store = std::atoi(e) ? 1 : 0;
  k.tma = (c.tile_mover == 2) ? 1 : 0;
  // experiment knobs; setting any of the pair knobs pins the scheduler to that setting (no plan portfolio)
  if (const char* e = std::getenv("QCB_PAIR_ROUNDS")) { k.pair_rounds = std::atoi(e) ? 1 : 0; k.plan_portfolio = 0; }
  if (const char* e = std::getenv("QCB_PAIR_COST_Q")) { k.pair_cost_q = std::max(4, std::atoi(e)); k.plan_portfolio = 0; }
  if (const char* e = std::getenv("QCB_PAIR_SEARCH")) { k.pair_search = std::max(1, std::atoi(e)); k.plan_portfolio = 0; }
  if (const char* e = std::getenv("QCB_PAIR_EFF_PCT")) { k.pair_eff_pct = std::atoi(e); k.plan_portfolio = 0; }
  if (const char* e = std::getenv("QCB_PLAN_PORTFOLIO")) k.plan_portfolio = std::atoi(e) ? 1 : 0;
  if (const char* e = std::getenv("QCB_FAR_PHASE")) k.far_phase = std::atoi(e) ? 1 : 0;
  if (const char* e = std::getenv("QCB_THIN_DEFER")) k.thin_defer = std::atoi(e);            // experiment knob
  if (const char* e = std::getenv("QCB_WINDOW_SEARCH")) k.window_search = std::atoi(e);
  if (const char* e = std::getenv("QCB_ROUND_YIELD_PCT")) k.round_yield_pct = std::atoi(e);   // 0 = greedy tiles / rounds only
  return k;
}

// ------------------------------------------------------------------ lowering (domain/circuit.clj:952-1072)
static bool qubit_ok(int q, int n) { return q >= 0 && q < n; }

int lower_ops(const Config& cfg, const qcb_op* ops, uint64_t n_ops, std::vector<Gate>& out, std::string& err) {
  const int n = cfg.n_total;
  auto bit = [&](int q) { return n - 1 - q; };
  auto fail = [&](int code, const std::string& m, uint64_t k) {
    err = m + " (op " + std::to_string(k) + ")";
    return code;
  };
  for (uint64_t k = 0; k < n_ops; ++k) {
    const qcb_op& op = ops[k];
    Gate g;
    g.src_op = (int)k;
    const int q0 = op.q[0], q1 = op.q[1], q2 = op.q[2];
    auto need1 = [&]() { return qubit_ok(q0, n); };
    auto need2 = [&]() { return qubit_ok(q0, n) && qubit_ok(q1, n) && q0 != q1; };
    auto need3 = [&]() { return need2() && qubit_ok(q2, n) && q2 != q0 && q2 != q1; };
    auto diag1 = [&](cplx ph) {  // multiply the b=1 half: Z,S,T,phase (gate.clj:38-137 through expand)
      g.kind = G_DMASK; g.dmask = g.dval = 1ULL << bit(q0); g.m[0] = ph; g.frac = 0.5;
    };
    switch (op.kind) {
      case QCB_OP_I:
        if (cfg.strict) return fail(QCB_ERR_UNSUPPORTED, "Unknown gate type :i", k);   // circuit.clj:1072
        if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k);
        continue;
      case QCB_OP_X: case QCB_OP_Y: case QCB_OP_H: case QCB_OP_RX: case QCB_OP_RY:
        if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k);
        g.kind = G_MAT1; g.t0 = bit(q0); gate_matrix(op.kind, op.angle, g.m); g.frac = 1.0;
        out.push_back(g); break;
      case QCB_OP_Z: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({-1, 0}); out.push_back(g); break;
      case QCB_OP_S: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({0, 1}); out.push_back(g); break;
      case QCB_OP_SDG: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({0, -1}); out.push_back(g); break;
      case QCB_OP_T: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({std::cos(kPi / 4), std::sin(kPi / 4)}); out.push_back(g); break;
      case QCB_OP_TDG: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({std::cos(kPi / -4), std::sin(kPi / -4)}); out.push_back(g); break;
      case QCB_OP_PHASE: if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k); diag1({std::cos(op.angle), std::sin(op.angle)}); out.push_back(g); break;
      case QCB_OP_RZ: {
        if (!need1()) return fail(QCB_ERR_INVALID, "bad target", k);
        cplx U[4]; gate_matrix(QCB_OP_RZ, op.angle, U);
        g.kind = G_DTAB1; g.t0 = bit(q0); g.m[0] = U[0]; g.m[1] = U[3]; g.frac = 1.0;
        out.push_back(g); break; }
      case QCB_OP_CNOT:
        if (!need2()) return fail(QCB_ERR_INVALID, "CNOT requires both control and target qubits", k);
        g.kind = G_MAT1; g.t0 = bit(q1); g.cmask = 1ULL << bit(q0); gate_matrix(QCB_OP_X, 0, g.m); g.frac = 0.5;
        out.push_back(g); break;
      case QCB_OP_CZ: case QCB_OP_RYDBERG_CZ:
        if (!need2()) return fail(QCB_ERR_INVALID, "CZ requires both control and target qubits", k);
        g.kind = G_DMASK; g.dmask = g.dval = (1ULL << bit(q0)) | (1ULL << bit(q1)); g.m[0] = {-1, 0}; g.frac = 0.25;
        out.push_back(g); break;
      case QCB_OP_CY: case QCB_OP_CRX: case QCB_OP_CRY: {
        if (op.kind == QCB_OP_CY && cfg.strict) return fail(QCB_ERR_UNSUPPORTED, "Unknown gate type :cy", k);
        if (!need2()) return fail(QCB_ERR_INVALID, "controlled gate requires control and target", k);
        cplx U[4]; gate_matrix(op.kind == QCB_OP_CY ? QCB_OP_Y : op.kind, op.angle, U);
        if (cfg.strict) std::swap(U[1], U[2]);        // reference applies U^T (gate.clj:473-483)
        g.kind = G_MAT1; g.t0 = bit(q1); g.cmask = 1ULL << bit(q0); std::memcpy(g.m, U, sizeof U); g.frac = 0.5;
        out.push_back(g); break; }
      case QCB_OP_CRZ: {
        if (!need2()) return fail(QCB_ERR_INVALID, "CRZ requires control, target qubits and angle", k);
        cplx U[4]; gate_matrix(QCB_OP_RZ, op.angle, U);
        g.kind = G_DTAB1; g.t0 = bit(q1); g.cmask = 1ULL << bit(q0); g.m[0] = U[0]; g.m[1] = U[3]; g.frac = 0.5;
        out.push_back(g); break; }
      case QCB_OP_SWAP: case QCB_OP_ISWAP: {
        if (!need2()) return fail(QCB_ERR_INVALID, "SWAP requires both qubit1 and qubit2 parameters", k);
        g.kind = G_SWAPP;
        // reference: operands are index-bit positions counted from the LSB (gate.clj:768-778, 823-833)
        int b0 = cfg.strict ? q0 : bit(q0), b1 = cfg.strict ? q1 : bit(q1);
        g.t0 = std::min(b0, b1); g.t1 = std::max(b0, b1);
        g.m[0] = (op.kind == QCB_OP_ISWAP) ? cplx{0, 1} : cplx{1, 0}; g.frac = 0.5;
        out.push_back(g); break; }
      case QCB_OP_TOFFOLI:
        if (!need3()) return fail(QCB_ERR_INVALID, "Toffoli requires control1, control2, and target parameters", k);
        g.kind = G_MAT1; g.t0 = bit(q2); g.cmask = (1ULL << bit(q0)) | (1ULL << bit(q1)); gate_matrix(QCB_OP_X, 0, g.m); g.frac = 0.25;
        out.push_back(g); break;
      case QCB_OP_FREDKIN: {
        if (!need3()) return fail(QCB_ERR_INVALID, "Fredkin requires control, target1, and target2 parameters", k);
        g.kind = G_SWAPP; int b1 = bit(q1), b2 = bit(q2);
        g.t0 = std::min(b1, b2); g.t1 = std::max(b1, b2); g.cmask = 1ULL << bit(q0); g.m[0] = {1, 0}; g.frac = 0.25;
        out.push_back(g); break; }
      case QCB_OP_RYDBERG_CPHASE:
        if (!need2()) return fail(QCB_ERR_INVALID, "Rydberg CPhase requires control, target qubits and phase angle", k);
        g.kind = G_DMASK; g.dmask = g.dval = (1ULL << bit(q0)) | (1ULL << bit(q1));
        g.m[0] = {std::cos(op.angle), std::sin(op.angle)}; g.frac = 0.25;
        out.push_back(g); break;
      case QCB_OP_RYDBERG_BLOCKADE: case QCB_OP_MCPHASE: {
        uint64_t mk = 0; int cnt = 0;
        for (int q = 0; q < 64; ++q)
          if ((op.mask >> q) & 1) { if (q >= n) return fail(QCB_ERR_INVALID, "qubit index out of range", k); mk |= 1ULL << bit(q); ++cnt; }
        if (cnt == 0) return fail(QCB_ERR_INVALID, "empty qubit set", k);
        g.dmask = mk; g.m[0] = {std::cos(op.angle), std::sin(op.angle)};
        if (op.kind == QCB_OP_RYDBERG_BLOCKADE) { g.kind = G_DPOP1; g.frac = (double)cnt / (double)(1ULL << cnt); }
        else { g.kind = G_DMASK; g.dval = mk; g.frac = 1.0 / (double)(1ULL << cnt); }
        out.push_back(g); break; }
      case QCB_OP_GLOBAL_H: case QCB_OP_GLOBAL_X: case QCB_OP_GLOBAL_Y: case QCB_OP_GLOBAL_Z:
      case QCB_OP_GLOBAL_RX: case QCB_OP_GLOBAL_RY: case QCB_OP_GLOBAL_RZ: {
        // gate.clj:1115-1253: the same 1q gate on every qubit, qubit 0 first; global-x/y/z = RX/RY/RZ(pi)
        int base = QCB_OP_H; double a = op.angle;
        if (op.kind == QCB_OP_GLOBAL_X) { base = QCB_OP_RX; a = kPi; }
        else if (op.kind == QCB_OP_GLOBAL_Y) { base = QCB_OP_RY; a = kPi; }
        else if (op.kind == QCB_OP_GLOBAL_Z) { base = QCB_OP_RZ; a = kPi; }
        else if (op.kind == QCB_OP_GLOBAL_RX) base = QCB_OP_RX;
        else if (op.kind == QCB_OP_GLOBAL_RY) base = QCB_OP_RY;
        else if (op.kind == QCB_OP_GLOBAL_RZ) base = QCB_OP_RZ;
        for (int q = 0; q < n; ++q) {
          Gate h; h.src_op = (int)k; h.frac = 1.0;
          cplx U[4]; gate_matrix(base, a, U);
          if (base == QCB_OP_RZ) { h.kind = G_DTAB1; h.t0 = bit(q); h.m[0] = U[0]; h.m[1] = U[3]; }
          else { h.kind = G_MAT1; h.t0 = bit(q); std::memcpy(h.m, U, sizeof U); }
          out.push_back(h);
        }
        break; }
      case QCB_OP_U1Q: case QCB_OP_CU1Q: {
        bool ctl = op.kind == QCB_OP_CU1Q;
        if (ctl ? !need2() : !need1()) return fail(QCB_ERR_INVALID, "bad qubits", k);
        g.kind = G_MAT1; g.t0 = bit(ctl ? q1 : q0); if (ctl) g.cmask = 1ULL << bit(q0);
        for (int i = 0; i < 4; ++i) g.m[i] = {op.mat[2 * i], op.mat[2 * i + 1]};
        g.frac = ctl ? 0.5 : 1.0;
        out.push_back(g); break; }
      case QCB_OP_U2Q: {
        if (!need2() || !op.ext) return fail(QCB_ERR_INVALID, "U2Q needs two qubits and a 4x4 matrix", k);
        g.kind = G_MAT2; int bh = bit(q0), bl = bit(q1);   // q0 = more significant basis bit of the 4x4
        cplx M[16];
        const double* ex = static_cast<const double*>(op.ext);
        for (int i = 0; i < 16; ++i) M[i] = {ex[2 * i], ex[2 * i + 1]};
        if (bh > bl) { g.t1 = bh; g.t0 = bl; std::memcpy(g.m, M, sizeof M); }
        else {  // reorder basis so that t1 (higher index bit) is the more significant basis bit
          g.t1 = bl; g.t0 = bh;
          auto sw = [](int i) { return ((i & 1) << 1) | ((i >> 1) & 1); };
          for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) g.m[sw(r) * 4 + sw(c)] = M[r * 4 + c];
        }
        g.frac = 1.0; out.push_back(g); break; }
      case QCB_OP_PHASE_ORACLE:
        if (n < 64 && (op.mask >> n) != 0) return fail(QCB_ERR_INVALID, "oracle index out of range", k);
        g.kind = G_DMASK; g.dmask = (n >= 64) ? ~0ULL : ((1ULL << n) - 1); g.dval = op.mask; g.m[0] = {-1, 0};
        g.frac = 1.0 / (double)(1ULL << std::min(n, 62)); out.push_back(g); break;
      case QCB_OP_GROVER_DIFFUSION:
        g.kind = G_REFLECT; g.frac = 1.5; out.push_back(g); break;   // read sweep + read/write sweep
      default:
        return fail(QCB_ERR_UNSUPPORTED, "Unknown gate type", k);
    }
  }
  return QCB_OK;
}

// ------------------------------------------------------------------ scheduling
static int gate_cost(const Gate& g) {
  switch (g.kind) {
    case G_MAT1: {
      const cplx* M = g.m;
      const bool perm = M[0].re == 0 && M[0].im == 0 && M[3].re == 0 && M[3].im == 0 && M[1].re == 1 && M[1].im == 0 && M[2].re == 1 && M[2].im == 0;
      if (perm) return 1;
      const bool real = M[0].im == 0 && M[1].im == 0 && M[2].im == 0 && M[3].im == 0;
      const bool ri = M[0].im == 0 && M[3].im == 0 && M[1].re == 0 && M[2].re == 0;
      int c = (real || ri) ? 4 : 8;
      return g.cmask ? (c + 1) / 2 : c;
    }
    case G_MAT2: return 32;
    case G_SWAPP: return 1;
    case G_DMASK: return (g.m[0].re == -1.0 && g.m[0].im == 0.0) ? 1 : 2;
    case G_DTAB1: return g.cmask ? 2 : 4;
    case G_DPOP1: return 3;
    default: return 4;
  }
}

struct Blocker {
  uint64_t x = 0, z = 0;   // bits used non-diagonally / diagonally by skipped gates
  bool conflicts(const Gate& g) const {
    uint64_t t = g.target_mask(), d = g.diag_mask();
    return (t & (x | z)) || (d & x);
  }
  void block(const Gate& g) { x |= g.target_mask(); z |= g.diag_mask(); }
};

static inline int popc(uint64_t v) { return __builtin_popcountll(v); }
constexpr size_t K3_FRAG_DOUBLES_HOST = 192;   // tile_core.h: K3_FRAG_DOUBLES

// Chunk bit (0..2) of the 128-byte shared-memory row that tile-local index bit p is XORed onto by the tile layout
// (tile_core.h: swz; c = run bits of the stage), or -1 when the bit does not move the bank group at all.
static int chunk_class(int p, int c) {
  if (p < 6) return p % 3;
  if (c >= 6) return -1;
  const int lo = c > 3 ? c - 3 : 0;
  return lo + (p - 6) % (3 - lo);
}

// choose lane positions so that the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups: three tile-local
// bits that are not slots, one per chunk class
static void choose_lanes(int m, int c, const std::vector<int>& slots, std::vector<int>& lanes) {
  lanes.clear();
  uint64_t used = 0;
  for (int s : slots) used |= 1ULL << s;
  int want = std::min(3, m - (int)slots.size());
  for (int k = 0; k < 3 && (int)lanes.size() < want; ++k) {
    int pick = -1;
    for (int p = 0; p < m; ++p)
      if (!((used >> p) & 1) && chunk_class(p, c) == k) { pick = p; break; }
    if (pick >= 0) { lanes.push_back(pick); used |= 1ULL << pick; }
  }
  for (int p = 0; p < m && (int)lanes.size() < want; ++p)
    if (!((used >> p) & 1)) { lanes.push_back(p); used |= 1ULL << p; }
}

struct SplitCond { uint32_t sel = 0, loc_mask = 0, loc_val = 0; uint64_t hi_mask = 0, hi_val = 0; };

// Split a condition (idx & mask) == val given in ext space into its slot / tile-local / tile-id+rank parts.
// `zero_slots`: slot indices that must be 0 in the patterns enumerated by sel (the op's own target slots).
static SplitCond split_cond(uint64_t mask, uint64_t val, const std::vector<int>& slot_pos, int m, uint32_t zero_slots) {
  SplitCond c;
  uint32_t ms = 0, vs = 0;
  for (int p = 0; p < 64; ++p) {
    if (!((mask >> p) & 1)) continue;
    const uint64_t vb = (val >> p) & 1;
    int j = -1;
    for (size_t k = 0; k < slot_pos.size(); ++k) if (slot_pos[k] == p) j = (int)k;
    if (j >= 0) { ms |= 1u << j; vs |= (uint32_t)vb << j; }
    else if (p < m) { c.loc_mask |= 1u << p; c.loc_val |= (uint32_t)vb << p; }
    else { c.hi_mask |= 1ULL << (p - m); c.hi_val |= vb << (p - m); }
  }
  const int r = (int)slot_pos.size();
  for (uint32_t s = 0; s < (1u << r); ++s)
    if ((s & zero_slots) == 0 && (s & ms) == vs) c.sel |= 1u << s;
  return c;
}

static uint64_t dbl_bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }

static size_t new_op(std::vector<uint64_t>& w, uint32_t kind, int j0, int j1, int nslots, const SplitCond& c) {
  size_t base = w.size();
  w.resize(base + (size_t)nslots * OP_WORDS, 0);
  w[base + 0] = (uint64_t)kind | ((uint64_t)(j0 < 0 ? 0 : j0) << 8) | ((uint64_t)(j1 < 0 ? 0 : j1) << 16) |
                ((uint64_t)nslots << 24) | ((uint64_t)c.sel << 32);
  w[base + 1] = (uint64_t)c.loc_mask | ((uint64_t)c.loc_val << 32);
  w[base + 2] = c.hi_mask;
  w[base + 3] = c.hi_val;
  return base;
}

static bool is_zero(cplx z) { return z.re == 0.0 && z.im == 0.0; }

static void put_gate_words(std::vector<uint64_t>& w, const Gate& g, const std::vector<int>& slot_pos, int m) {
  auto slot_of = [&](int pos) {
    for (size_t j = 0; j < slot_pos.size(); ++j) if (slot_pos[j] == pos) return (int)j;
    return -1;
  };
  switch (g.kind) {
    case G_MAT1: {
      const int j = slot_of(g.t0);
      SplitCond c = split_cond(g.cmask, g.cmask, slot_pos, m, 1u << j);
      const cplx* M = g.m;
      if (is_zero(M[0]) && is_zero(M[3]) && M[1].re == 1.0 && M[1].im == 0.0 && M[2].re == 1.0 && M[2].im == 0.0) {
        new_op(w, D_PERMX, j, -1, 1, c);
      } else if (M[0].im == 0.0 && M[1].im == 0.0 && M[2].im == 0.0 && M[3].im == 0.0) {
        size_t b = new_op(w, D_MAT1R, j, -1, 1, c);
        for (int i = 0; i < 4; ++i) w[b + 4 + i] = dbl_bits(M[i].re);
      } else if (M[0].im == 0.0 && M[3].im == 0.0 && M[1].re == 0.0 && M[2].re == 0.0) {
        size_t b = new_op(w, D_MAT1RI, j, -1, 1, c);
        w[b + 4] = dbl_bits(M[0].re); w[b + 5] = dbl_bits(M[1].im); w[b + 6] = dbl_bits(M[2].im); w[b + 7] = dbl_bits(M[3].re);
      } else {
        size_t b = new_op(w, D_MAT1, j, -1, 1, c);
        for (int i = 0; i < 4; ++i) { w[b + 4 + 2 * i] = dbl_bits(M[i].re); w[b + 5 + 2 * i] = dbl_bits(M[i].im); }
      }
      break;
    }
    case G_MAT2: {
      const int j0 = slot_of(g.t0), j1 = slot_of(g.t1);
      SplitCond c = split_cond(g.cmask, g.cmask, slot_pos, m, (1u << j0) | (1u << j1));
      size_t b = new_op(w, D_MAT2, j0, j1, 3, c);
      for (int i = 0; i < 16; ++i) { w[b + 4 + 2 * i] = dbl_bits(g.m[i].re); w[b + 5 + 2 * i] = dbl_bits(g.m[i].im); }
      break;
    }
    case G_SWAPP: {
      const int j0 = slot_of(g.t0), j1 = slot_of(g.t1);
      SplitCond c = split_cond(g.cmask, g.cmask, slot_pos, m, (1u << j0) | (1u << j1));
      size_t b = new_op(w, D_SWAPP, j0, j1, 1, c);
      w[b + 4] = dbl_bits(g.m[0].re); w[b + 5] = dbl_bits(g.m[0].im);
      break;
    }
    case G_DMASK: {
      SplitCond c = split_cond(g.dmask, g.dval, slot_pos, m, 0);
      if (g.m[0].re == -1.0 && g.m[0].im == 0.0) { new_op(w, D_DNEG, -1, -1, 1, c); break; }
      size_t b = new_op(w, D_DMASK, -1, -1, 1, c);
      w[b + 4] = dbl_bits(g.m[0].re); w[b + 5] = dbl_bits(g.m[0].im);
      break;
    }
    case G_DTAB1: {   // phase table on one bit = two masked phases (bit = 0 -> m[0], bit = 1 -> m[1])
      for (int bval = 0; bval < 2; ++bval) {
        const cplx ph = g.m[bval];
        if (ph.re == 1.0 && ph.im == 0.0) continue;
        const uint64_t mask = g.cmask | (1ULL << g.t0), val = g.cmask | ((uint64_t)bval << g.t0);
        SplitCond c = split_cond(mask, val, slot_pos, m, 0);
        if (ph.re == -1.0 && ph.im == 0.0) { new_op(w, D_DNEG, -1, -1, 1, c); continue; }
        size_t b = new_op(w, D_DMASK, -1, -1, 1, c);
        w[b + 4] = dbl_bits(ph.re); w[b + 5] = dbl_bits(ph.im);
      }
      break;
    }
    case G_DENSE: {
      const int r = (int)slot_pos.size(), dim = 1 << r;
      SplitCond c; c.sel = 0xff;
      const int nwords = 4 + 2 * dim * dim;
      const int nslots = (nwords + OP_WORDS - 1) / OP_WORDS;
      size_t b = new_op(w, D_DENSE, -1, -1, nslots, c);
      for (int i = 0; i < dim * dim; ++i) { w[b + 4 + 2 * i] = dbl_bits(g.dense[i].re); w[b + 5 + 2 * i] = dbl_bits(g.dense[i].im); }
      break;
    }
    case G_DPOP1: {
      SplitCond c; c.sel = 0xff;
      size_t b = new_op(w, D_DPOP1, -1, -1, 1, c);
      w[b + 4] = g.dmask; w[b + 6] = dbl_bits(g.m[0].re); w[b + 7] = dbl_bits(g.m[0].im);
      break;
    }
    default: {        // G_REFLECT -> affine a' = alpha a + beta with coefficients in device values slot dval
      SplitCond c; c.sel = 0xff;
      size_t b = new_op(w, D_AFFINE, -1, -1, 1, c);
      w[b + 4] = g.dval;
      break;
    }
  }
}

// translate a gate from physical bit space into the ext space of a stage
static Gate to_ext(const Gate& g, const std::vector<int>& ext_of_phys) {
  Gate e = g;
  auto mp = [&](uint64_t mask) {
    uint64_t r = 0;
    for (int b = 0; b < 64; ++b) if ((mask >> b) & 1) r |= 1ULL << ext_of_phys[b];
    return r;
  };
  if (g.t0 >= 0) e.t0 = ext_of_phys[g.t0];
  if (g.t1 >= 0) e.t1 = ext_of_phys[g.t1];
  e.cmask = mp(g.cmask);
  if (g.kind == G_DMASK || g.kind == G_DPOP1) { e.dmask = mp(g.dmask); e.dval = mp(g.dval); }
  if (g.kind == G_MAT2 && e.t0 > e.t1) {   // keep t1 as the higher ext position: permute the basis
    std::swap(e.t0, e.t1);
    auto sw = [](int i) { return ((i & 1) << 1) | ((i >> 1) & 1); };
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) e.m[sw(r) * 4 + sw(c)] = g.m[r * 4 + c];
  }
  if (g.kind == G_SWAPP && e.t0 > e.t1) std::swap(e.t0, e.t1);
  return e;
}

static void encode_stage(const Config& cfg, Stage& st, std::vector<uint64_t>& words) {
  {
    size_t need = STAGE_WORDS + st.rounds.size() * ROUND_WORDS;
    for (const Round& rd : st.rounds) need += rd.frag.size() + rd.far.size() + rd.far2.size() + rd.farpre.size() + rd.farpre2.size() + 64 * OP_WORDS;
    if (words.capacity() < words.size() + need) words.reserve(std::max(words.capacity() * 2, words.size() + need));
  }
  size_t base = words.size();
  words.resize(base + STAGE_WORDS, 0);
  const int m = st.m;
  words[base + 0] = (uint64_t)cfg.n_local;
  words[base + 1] = (uint64_t)m;
  words[base + 2] = (uint64_t)st.L;
  words[base + 3] = (uint64_t)st.rounds.size();
  words[base + 5] = (uint64_t)cfg.rank << (cfg.n_local - m);
  words[base + 6] = st.skip_mask;
  words[base + 7] = st.skip_val;
  for (int k = 0; k < m; ++k) words[base + 8 + k] = (uint64_t)st.tile_pos[k];
  // runs of non-tile positions
  uint64_t tmask = 0;
  for (int p : st.tile_pos) tmask |= 1ULL << p;
  int nruns = 0;
  for (int p = 0; p < cfg.n_local;) {
    if ((tmask >> p) & 1) { ++p; continue; }
    int s = p;
    while (p < cfg.n_local && !((tmask >> p) & 1)) ++p;
    words[base + 24 + nruns] = (uint64_t)s | ((uint64_t)(p - s) << 8);
    ++nruns;
  }
  words[base + 4] = (uint64_t)nruns;
  words[base + 41] = st.flags;
  words[base + 43] = (uint64_t)st.layout_c;
  // rounds: descriptors, then interpreter op slots, then (not copied to shared memory) tensor-core matrices
  size_t rbase = words.size();
  words.resize(rbase + st.rounds.size() * ROUND_WORDS, 0);
  for (size_t r = 0; r < st.rounds.size(); ++r) {
    Round& rd = st.rounds[r];
    size_t rb = rbase + r * ROUND_WORDS;
    words[rb + 0] = rd.slot_pos.size();
    for (size_t j = 0; j < rd.slot_pos.size(); ++j) words[rb + 4 + j] = (uint64_t)rd.slot_pos[j];
    if (rd.dmma) {
      words[rb + 17] = rd.pair ? 3 : (rd.k3 ? 2 : 1);
      if (rd.pair) words[rb + 36] = (uint64_t)rd.mmap2[0] | ((uint64_t)rd.mmap2[1] << 4) | ((uint64_t)rd.mmap2[2] << 8);
      words[rb + 18] = rd.grp_pos.size();
      for (size_t j = 0; j < rd.grp_pos.size() && j < 10; ++j) words[rb + 19 + j] = (uint64_t)rd.grp_pos[j];
      words[rb + 29] = rd.cond_pos.size();
      for (size_t j = 0; j < rd.cond_pos.size(); ++j) words[rb + 30 + j] = (uint64_t)rd.cond_pos[j];
      words[rb + 34] = rd.k3 ? ((uint64_t)rd.kmap[0] | ((uint64_t)rd.kmap[1] << 4) | ((uint64_t)rd.kmap[2] << 8)) : (uint64_t)rd.j_load;
      words[rb + 35] = rd.k3 ? ((uint64_t)rd.mmap[0] | ((uint64_t)rd.mmap[1] << 4) | ((uint64_t)rd.mmap[2] << 8)) : (uint64_t)rd.j_store;
      continue;
    }
    std::vector<int> lanes;
    choose_lanes(m, st.layout_c, rd.slot_pos, lanes);
    size_t ob = words.size();
    for (const Gate& g : rd.gates) put_gate_words(words, g, rd.slot_pos, m);
    rb = rbase + r * ROUND_WORDS;
    words[rb + 1] = (words.size() - ob) / OP_WORDS;
    words[rb + 2] = ob - base;
    words[rb + 3] = lanes.size();
    for (size_t j = 0; j < lanes.size(); ++j) words[rb + 7 + j] = (uint64_t)lanes[j];
    std::vector<int> ins(rd.slot_pos);
    ins.insert(ins.end(), lanes.begin(), lanes.end());
    std::sort(ins.begin(), ins.end());
    words[rb + 10] = ins.size();
    for (size_t j = 0; j < ins.size(); ++j) words[rb + 11 + j] = (uint64_t)ins[j];
  }
  words[base + 42] = words.size() - base;          // descriptor part (copied to shared memory by the kernel)
  for (size_t r = 0; r < st.rounds.size(); ++r) {
    Round& rd = st.rounds[r];
    if (!rd.dmma) continue;
    size_t rb = rbase + r * ROUND_WORDS;
    words[rb + 2] = words.size() - base;
    {
      const size_t at = words.size();
      words.resize(at + rd.frag.size());
      std::memcpy(words.data() + at, rd.frag.data(), rd.frag.size() * sizeof(double));      // a double's bits are the word
    }
    uint64_t off[4] = {0, 0, 0, 0};
    const std::vector<uint64_t>* tabs[4] = {&rd.far, &rd.far2, &rd.farpre, &rd.farpre2};
    for (int t = 0; t < 4; ++t) if (!tabs[t]->empty()) { off[t] = words.size() - base; for (uint64_t w : *tabs[t]) words.push_back(w); }
    words[rb + 37] = off[0] | (off[2] << 32);              // first block: after (row) | before (column) << 32
    words[rb + 38] = off[1] | (off[3] << 32);              // second block
    words[rb + 39] = (uint64_t)(rd.far.size() / 5) | ((uint64_t)(rd.far2.size() / 5) << 8) | ((uint64_t)(rd.farpre.size() / 5) << 16) |
                     ((uint64_t)(rd.farpre2.size() / 5) << 24);
  }
  words[base + 40] = words.size() - base;
}

// ---- dense fusion inside a round (north_star: "merges runs of gates on at most k qubits into one dense
// 2^k x 2^k unitary applied from shared-memory tiles").  A gate is *pure* for a round when every bit it reads or
// writes is one of the round's slot bits; a run of pure gates is multiplied on the host into one matrix.
static bool gate_is_pure(const Gate& g, uint64_t slot_mask) {
  switch (g.kind) {
    case G_MAT1: case G_MAT2: case G_SWAPP: return ((g.target_mask() | g.cmask) & ~slot_mask) == 0;
    case G_DMASK: return (g.dmask & ~slot_mask) == 0;
    case G_DTAB1: return ((g.cmask | (1ULL << g.t0)) & ~slot_mask) == 0;
    default: return false;
  }
}

static inline cplx cm(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cplx ca(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }

// apply gate g (ext space) to a vector over the 2^r slot patterns.  Bits of g that are not slot bits are read
// from `fixed` (an ext-space bit assignment): with them fixed, every supported gate acts linearly on the slots.
static void small_apply(const Gate& g, const std::vector<int>& slot_pos, std::vector<cplx>& v, uint64_t fixed = 0) {
  const int r = (int)slot_pos.size(), dim = 1 << r;
  uint64_t slot_mask = 0;
  for (int p : slot_pos) slot_mask |= 1ULL << p;
  auto sbit = [&](int pos) { for (int j = 0; j < r; ++j) if (slot_pos[j] == pos) return j; return -1; };
  auto smask = [&](uint64_t m) { uint32_t o = 0; for (int j = 0; j < r; ++j) if ((m >> slot_pos[j]) & 1) o |= 1u << j; return o; };
  auto outside_ok = [&](uint64_t mask, uint64_t val) { const uint64_t mo = mask & ~slot_mask; return (fixed & mo) == (val & mo); };
  cplx o[1 << MAX_SLOT_BITS];                       // no heap traffic: this runs ~10^5 times per plan
  for (int s = 0; s < dim; ++s) o[s] = v[s];
  switch (g.kind) {
    case G_MAT1: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j = sbit(g.t0); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if ((s >> j) & 1) continue;
        if ((s & c) != c) continue;
        const int s1 = s | (1 << j);
        o[s] = ca(cm(g.m[0], v[s]), cm(g.m[1], v[s1]));
        o[s1] = ca(cm(g.m[2], v[s]), cm(g.m[3], v[s1]));
      }
      break;
    }
    case G_MAT2: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j0 = sbit(g.t0), j1 = sbit(g.t1); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if (((s >> j0) & 1) || ((s >> j1) & 1) || (s & c) != c) continue;
        const int id[4] = {s, s | (1 << j0), s | (1 << j1), s | (1 << j0) | (1 << j1)};
        for (int row = 0; row < 4; ++row) {
          cplx acc{0, 0};
          for (int col = 0; col < 4; ++col) acc = ca(acc, cm(g.m[row * 4 + col], v[id[col]]));
          o[id[row]] = acc;
        }
      }
      break;
    }
    case G_SWAPP: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j0 = sbit(g.t0), j1 = sbit(g.t1); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if (((s >> j0) & 1) || ((s >> j1) & 1) || (s & c) != c) continue;
        const int u = s | (1 << j0), w = s | (1 << j1);
        o[u] = cm(g.m[0], v[w]); o[w] = cm(g.m[0], v[u]);
      }
      break;
    }
    case G_DMASK: {
      if (!outside_ok(g.dmask, g.dval)) break;
      const uint32_t mk = smask(g.dmask), vl = smask(g.dval);
      for (int s = 0; s < dim; ++s) if ((s & mk) == vl) o[s] = cm(v[s], g.m[0]);
      break;
    }
    case G_DTAB1: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j = sbit(g.t0); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if ((s & c) != c) continue;
        const int tb = (j >= 0) ? ((s >> j) & 1) : (int)((fixed >> g.t0) & 1);
        o[s] = cm(v[s], g.m[tb]);
      }
      break;
    }
    case G_DPOP1: {
      const int outside_ones = popc(fixed & g.dmask & ~slot_mask);
      const uint32_t mk = smask(g.dmask);
      for (int s = 0; s < dim; ++s) if (outside_ones + popc((uint64_t)(s & mk)) == 1) o[s] = cm(v[s], g.m[0]);
      break;
    }
    default: break;
  }
  for (int s = 0; s < dim; ++s) v[s] = o[s];
}

// The same map applied to `ncol` vectors at once (M[pattern][column]): what build_dmma_round needs for the 8 basis columns of
// a round's dense block.  Per column the arithmetic and its order are exactly small_apply's.
static void small_apply_cols(const Gate& g, const std::vector<int>& slot_pos, cplx (*M)[8], int ncol, uint64_t fixed) {
  const int r = (int)slot_pos.size(), dim = 1 << r;
  uint64_t slot_mask = 0;
  for (int p : slot_pos) slot_mask |= 1ULL << p;
  auto sbit = [&](int pos) { for (int j = 0; j < r; ++j) if (slot_pos[j] == pos) return j; return -1; };
  auto smask = [&](uint64_t m) { uint32_t o = 0; for (int j = 0; j < r; ++j) if ((m >> slot_pos[j]) & 1) o |= 1u << j; return o; };
  auto outside_ok = [&](uint64_t mask, uint64_t val) { const uint64_t mo = mask & ~slot_mask; return (fixed & mo) == (val & mo); };
  switch (g.kind) {
    case G_MAT1: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j = sbit(g.t0); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if (((s >> j) & 1) || (s & c) != c) continue;
        const int s1 = s | (1 << j);
        for (int col = 0; col < ncol; ++col) {
          const cplx a = M[s][col], b = M[s1][col];
          M[s][col] = ca(cm(g.m[0], a), cm(g.m[1], b));
          M[s1][col] = ca(cm(g.m[2], a), cm(g.m[3], b));
        }
      }
      break;
    }
    case G_MAT2: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j0 = sbit(g.t0), j1 = sbit(g.t1); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if (((s >> j0) & 1) || ((s >> j1) & 1) || (s & c) != c) continue;
        const int id[4] = {s, s | (1 << j0), s | (1 << j1), s | (1 << j0) | (1 << j1)};
        for (int col = 0; col < ncol; ++col) {
          const cplx in[4] = {M[id[0]][col], M[id[1]][col], M[id[2]][col], M[id[3]][col]};
          for (int row = 0; row < 4; ++row) {
            cplx acc{0, 0};
            for (int k = 0; k < 4; ++k) acc = ca(acc, cm(g.m[row * 4 + k], in[k]));
            M[id[row]][col] = acc;
          }
        }
      }
      break;
    }
    case G_SWAPP: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j0 = sbit(g.t0), j1 = sbit(g.t1); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if (((s >> j0) & 1) || ((s >> j1) & 1) || (s & c) != c) continue;
        const int u = s | (1 << j0), w = s | (1 << j1);
        for (int col = 0; col < ncol; ++col) {
          const cplx a = M[u][col], b = M[w][col];
          M[u][col] = cm(g.m[0], b); M[w][col] = cm(g.m[0], a);
        }
      }
      break;
    }
    case G_DMASK: {
      if (!outside_ok(g.dmask, g.dval)) break;
      const uint32_t mk = smask(g.dmask), vl = smask(g.dval);
      for (int s = 0; s < dim; ++s) if ((s & mk) == vl) for (int col = 0; col < ncol; ++col) M[s][col] = cm(M[s][col], g.m[0]);
      break;
    }
    case G_DTAB1: {
      if (!outside_ok(g.cmask, g.cmask)) break;
      const int j = sbit(g.t0); const uint32_t c = smask(g.cmask);
      for (int s = 0; s < dim; ++s) {
        if ((s & c) != c) continue;
        const int tb = (j >= 0) ? ((s >> j) & 1) : (int)((fixed >> g.t0) & 1);
        for (int col = 0; col < ncol; ++col) M[s][col] = cm(M[s][col], g.m[tb]);
      }
      break;
    }
    case G_DPOP1: {
      const int outside_ones = popc(fixed & g.dmask & ~slot_mask);
      const uint32_t mk = smask(g.dmask);
      for (int s = 0; s < dim; ++s)
        if (outside_ones + popc((uint64_t)(s & mk)) == 1) for (int col = 0; col < ncol; ++col) M[s][col] = cm(M[s][col], g.m[0]);
      break;
    }
    default: break;
  }
}

// bits of a gate that are read or written (ext space)
static uint64_t gate_bits(const Gate& g) {
  uint64_t b = g.target_mask() | g.diag_mask();
  return b;
}

// ---- far phases.  A diagonal gate on exactly two bits (CRZ, CZ, controlled phase after lowering) one of which is a slot of the
// round and the other lies OUTSIDE the tile (a tile-id or rank bit) needs no condition bit: for a given tile the far bit is a
// constant, so the gate is a diagonal on one slot whose angle is a constant of the tile.  The product of all such gates of a
// round is  D(tile) = e^{i gamma} prod_j diag(e^{-i phi_j}, e^{+i phi_j})_slot j,  gamma / phi_j = (constants, folded into the
// matrices here) + sums over the far bits that are set - a ROW scaling of the round's 8x8 block as long as no later gate of the
// round acts non-diagonally on that slot.  The kernel applies it per tile and pass to the A fragments (kernels.cu: far_factor).
static bool far_kind(const Gate& g) {
  if (g.kind == G_DTAB1) return popc(g.cmask) == 1 && g.t0 >= 0 && !((g.cmask >> g.t0) & 1);
  if (g.kind == G_DMASK) return popc(g.dmask) == 2;
  return false;
}
// flags[i] = gate i of `gates` rides as a far phase of a round with the CHOSEN slot bits slot_mask (tile of m bits)
static void classify_far(const Config& cfg, const std::vector<Gate>& gates, uint64_t slot_mask, int m, std::vector<char>& flags) {
  flags.assign(gates.size(), 0);
  if (!cfg.far_phase || cfg.mma_form != 0) return;
  const uint64_t tile_mask = (1ULL << m) - 1ULL;
  uint64_t later_targets = 0;
  for (size_t i = gates.size(); i-- > 0;) {
    const Gate& g = gates[i];
    if (far_kind(g)) {
      const uint64_t b = gate_bits(g), l = b & slot_mask, h = b & ~tile_mask;
      if (popc(b) == 2 && popc(l) == 1 && popc(h) == 1 && !(later_targets & l)) flags[i] = 1;      // after the block: row scaling
    }
    later_targets |= g.target_mask();
  }
  // the same BEFORE the block (column scaling) when no earlier gate of the round acts non-diagonally on the slot
  uint64_t earlier_targets = 0;
  for (size_t i = 0; i < gates.size(); ++i) {
    const Gate& g = gates[i];
    if (!flags[i] && far_kind(g)) {
      const uint64_t b = gate_bits(g), l = b & slot_mask, h = b & ~tile_mask;
      if (popc(b) == 2 && popc(l) == 1 && popc(h) == 1 && !(earlier_targets & l)) flags[i] = 2;
    }
    earlier_targets |= g.target_mask();
  }
}
// condition bits of a round: every bit a gate touches that is not a slot, far phases excepted
static uint64_t round_cond_bits(const std::vector<Gate>& gates, const std::vector<char>& far, uint64_t slot_mask) {
  uint64_t cond = 0;
  for (size_t i = 0; i < gates.size(); ++i) if (!far[i]) cond |= gate_bits(gates[i]) & ~slot_mask;
  return cond;
}
// Far-phase table of one block: slots = final slot positions (pattern bit j <-> slots[j]).  Entry = {far position - m, gamma,
// phi_0, phi_1, phi_2}.  With D0 / D1 the gate's diagonal on its slot for the far bit = 0 / 1: D0 goes into the matrices
// (far bit forced to 0 by the caller), D1 / D0 = e^{i g} diag(e^{-i f}, e^{+i f}) into the table.
static void build_far_table(const std::vector<Gate>& gates, const std::vector<char>& far, const std::vector<int>& slots, int m,
                            std::vector<uint64_t>& out, char which = 1) {
  out.clear();
  std::vector<int> pos;
  std::vector<double> val;                       // 4 per entry
  const uint64_t tile_mask = (1ULL << m) - 1ULL;
  for (size_t i = 0; i < gates.size(); ++i) {
    if (far[i] != which) continue;
    const Gate& g = gates[i];
    const uint64_t b = gate_bits(g), hm = b & ~tile_mask;
    const int h = 63 - __builtin_clzll(hm);
    int j = -1, lp = -1;
    for (size_t k = 0; k < slots.size(); ++k) if ((b >> slots[k]) & 1) { j = (int)k; lp = slots[k]; }
    std::vector<int> one = {lp};
    std::vector<cplx> v0 = {cplx{1, 0}, cplx{1, 0}}, v1 = v0;
    small_apply(g, one, v0, 0);
    small_apply(g, one, v1, hm);
    auto ratio_arg = [](cplx a1, cplx a0) { return std::atan2(a1.im * a0.re - a1.re * a0.im, a1.re * a0.re + a1.im * a0.im); };   // arg(a1 / a0)
    const double ra = ratio_arg(v1[0], v0[0]), rb = ratio_arg(v1[1], v0[1]);
    size_t e = 0;
    while (e < pos.size() && pos[e] != h - m) ++e;
    if (e == pos.size()) { pos.push_back(h - m); val.insert(val.end(), 4, 0.0); }
    val[4 * e] += 0.5 * (ra + rb);
    val[4 * e + 1 + j] += 0.5 * (rb - ra);
  }
  for (size_t e = 0; e < pos.size(); ++e) {
    out.push_back((uint64_t)pos[e]);
    for (int k = 0; k < 4; ++k) out.push_back(dbl_bits(val[4 * e + k]));
  }
}

// ---- tensor-core round: the whole round as 2^k dense 16x16 real matrices in mma.m16n8k16 A-fragment order
static bool dmma_eligible(const Config& cfg, const Stage& st, const Round& rd) {
  if (!cfg.dense_mma || st.m < 6 || rd.slot_pos.size() > 3) return false;
  uint64_t slot_mask = 0;
  for (int p : rd.slot_pos) slot_mask |= 1ULL << p;
  for (const Gate& g : rd.gates) if (g.kind == G_REFLECT || g.kind == G_DENSE) return false;
  std::vector<char> far;
  classify_far(cfg, rd.gates, slot_mask, st.m, far);
  const uint64_t cond = round_cond_bits(rd.gates, far, slot_mask);
  // padding the slot set to 3 needs free tile-local bits; lanes need 3 more
  const int free_bits = st.m - (int)rd.slot_pos.size() - popc(cond & ((1ULL << st.m) - 1ULL));
  if (free_bits < (3 - (int)rd.slot_pos.size()) + 3) return false;
  return popc(cond) <= MAX_COND_BITS;
}

static void build_dmma_round(const Config& cfg, const Stage& st, Round& rd) {
  const int m = st.m;
  uint64_t slot_mask = 0;
  for (int p : rd.slot_pos) slot_mask |= 1ULL << p;
  // condition bits = every bit a gate touches that is not a slot bit
  uint64_t cond = 0;
  for (const Gate& g : rd.gates) cond |= gate_bits(g) & ~slot_mask;
  rd.cond_pos.clear();
  for (int p = 0; p < 64; ++p) if ((cond >> p) & 1) rd.cond_pos.push_back(p);
  const uint64_t tile_mask = (1ULL << m) - 1ULL;
  // pad the slot set to 3 bits with unused tile-local bits (identity on them), highest first
  for (int p = m - 1; p >= 0 && rd.slot_pos.size() < 3; --p)
    if (!((slot_mask >> p) & 1) && !((cond >> p) & 1)) { rd.slot_pos.push_back(p); slot_mask |= 1ULL << p; }
  std::sort(rd.slot_pos.begin(), rd.slot_pos.end());
  // lane bits: 3 tile-local bits that are neither slots nor conditions, one per chunk class of the tile layout
  // (chunk_class); then choose which slot bit rides on the half-warp's thread index for loads (j_load) and stores
  // (j_store) so that 8-byte accesses are conflict-free
  const int rb = st.layout_c;
  const uint64_t busy = slot_mask | (cond & tile_mask);
  std::vector<int> lanes;
  for (int cl = 0; cl < 3; ++cl)
    for (int p = 0; p < m; ++p)
      if (!((busy >> p) & 1) && chunk_class(p, rb) == cl) { lanes.push_back(p); break; }
  uint64_t used = busy;
  for (int p : lanes) used |= 1ULL << p;
  for (int p = 0; p < m && lanes.size() < 3; ++p) if (!((used >> p) & 1)) { lanes.push_back(p); used |= 1ULL << p; }
  // exhaustive search over the candidate lane bits (every free tile-local bit), their order and the slot choices:
  // loads vary (lanes[0], lanes[1], slot j_load) inside a half-warp, stores vary (lanes[1], lanes[2], slot j_store);
  // each triple wants three distinct chunk classes
  int jl = 0, js = 0;
  {
    std::vector<int> cand;
    for (int p = 0; p < m; ++p) if (!((busy >> p) & 1)) cand.push_back(p);
    auto distinct3 = [&](int a, int b, int d) {
      const int ca = chunk_class(a, rb), cb = chunk_class(b, rb), cd = chunk_class(d, rb);
      return ca >= 0 && cb >= 0 && cd >= 0 && ca != cb && cb != cd && ca != cd;
    };
    int best = -1;
    std::vector<int> bestl = lanes;
    const int nc = (int)cand.size();
    for (int i0 = 0; i0 < nc && best < 2; ++i0) for (int i1 = 0; i1 < nc && best < 2; ++i1) for (int i2 = 0; i2 < nc; ++i2) {
      if (i0 == i1 || i1 == i2 || i0 == i2) continue;
      const int l0 = cand[i0], l1 = cand[i1], l2 = cand[i2];
      int bl = -1, bs = -1;
      for (int a = 0; a < 3 && bl < 0; ++a) if (distinct3(l0, l1, rd.slot_pos[a])) bl = a;
      for (int b = 0; b < 3 && bs < 0; ++b) if (distinct3(l1, l2, rd.slot_pos[b])) bs = b;
      const int score = (bl >= 0) + (bs >= 0);
      if (score > best) { best = score; jl = bl >= 0 ? bl : 0; js = bs >= 0 ? bs : 0; bestl = {l0, l1, l2}; }
      if (best == 2) break;
    }
    if (bestl.size() == 3) lanes = bestl;
  }
  rd.j_load = jl; rd.j_store = js;
  // group-index bit order: lane bits, then free bits ascending, then tile-local condition bits (so that the
  // batches of one warp share the variant)
  rd.grp_pos = lanes;
  uint64_t lane_mask = 0;
  for (int p : lanes) lane_mask |= 1ULL << p;
  for (int p = 0; p < m; ++p) if (!(((slot_mask | lane_mask | cond) >> p) & 1)) rd.grp_pos.push_back(p);
  for (int p = 0; p < m; ++p) if (((cond >> p) & 1) && !((slot_mask >> p) & 1)) rd.grp_pos.push_back(p);
  // matrices
  const int k = (int)rd.cond_pos.size();
  const size_t nvar = (size_t)1 << k;
  rd.frag.assign(nvar * 256, 0.0);
  // k-index / m-index -> (slot pattern, component): bit0 = component, bit1 = slot j (j_load / j_store),
  // bits 2,3 = the two remaining slots in ascending slot order
  auto idx_map = [&](int idx, int jsel, int& pattern, int& comp) {
    comp = idx & 1;
    int rem[2], nr = 0;
    for (int j = 0; j < 3; ++j) if (j != jsel) rem[nr++] = j;
    pattern = (((idx >> 1) & 1) << jsel) | (((idx >> 2) & 1) << rem[0]) | (((idx >> 3) & 1) << rem[1]);
  };
  for (size_t var = 0; var < nvar; ++var) {
    uint64_t fixed = 0;
    for (int j = 0; j < k; ++j) if ((var >> j) & 1) fixed |= 1ULL << rd.cond_pos[j];
    cplx M[8][8];
    for (int row = 0; row < 8; ++row) for (int col = 0; col < 8; ++col) M[row][col] = cplx{row == col ? 1.0 : 0.0, 0.0};
    for (const Gate& g : rd.gates) small_apply_cols(g, rd.slot_pos, M, 8, fixed);
    double W[16][16];
    for (int mi = 0; mi < 16; ++mi) for (int ki = 0; ki < 16; ++ki) {
      int pr, cr, pc, cc;
      idx_map(mi, rd.j_store, pr, cr);
      idx_map(ki, rd.j_load, pc, cc);
      const cplx z = M[pr][pc];
      // (re_out, im_out) = [[zr, -zi],[zi, zr]] (re_in, im_in)
      W[mi][ki] = (cr == 0) ? (cc == 0 ? z.re : -z.im) : (cc == 0 ? z.im : z.re);
    }
    for (int i = 0; i < 8; ++i) for (int lane = 0; lane < 32; ++lane) {
      const int mrow = lane / 4 + 8 * (i & 1), kcol = lane % 4 + 4 * (i >> 1);
      rd.frag[var * 256 + (size_t)i * 32 + lane] = W[mrow][kcol];
    }
  }
  rd.dmma = true;
  (void)cfg;
}

// The same round in the three-product form (tile_core.h: K3Ctx): 2^k variants of three 8x8 real matrices P = Mr + Mi,
// N = -Mi, R = Mr in mma.m8n8k4 A-fragment order.  Shared-memory accesses are 16 bytes per lane, served per quarter-warp:
// a load quarter varies k-index bits 0, 1 (two slot bits) and group bit 0; a store quarter varies m-index bit 0 (one slot
// bit) and group bits 1, 2 - each triple wants three distinct chunk classes of the tile layout (chunk_class).
static void build_k3_round(const Config& cfg, const Stage& st, Round& rd) {
  const int m = st.m;
  uint64_t slot_mask = 0;
  for (int p : rd.slot_pos) slot_mask |= 1ULL << p;
  std::vector<char> far;
  classify_far(cfg, rd.gates, slot_mask, m, far);
  const uint64_t cond = round_cond_bits(rd.gates, far, slot_mask);
  const uint64_t not_tile = ~((1ULL << m) - 1ULL);
  rd.cond_pos.clear();
  for (int p = 0; p < 64; ++p) if ((cond >> p) & 1) rd.cond_pos.push_back(p);
  const uint64_t tile_mask = (1ULL << m) - 1ULL;
  for (int p = m - 1; p >= 0 && rd.slot_pos.size() < 3; --p)
    if (!((slot_mask >> p) & 1) && !((cond >> p) & 1)) { rd.slot_pos.push_back(p); slot_mask |= 1ULL << p; }
  std::sort(rd.slot_pos.begin(), rd.slot_pos.end());
  const int rb = st.layout_c;
  const uint64_t busy = slot_mask | (cond & tile_mask);
  std::vector<int> cand;
  for (int p = 0; p < m; ++p) if (!((busy >> p) & 1)) cand.push_back(p);
  auto distinct3 = [&](int a, int b, int d) {
    const int ca = chunk_class(a, rb), cb = chunk_class(b, rb), cd = chunk_class(d, rb);
    return ca >= 0 && cb >= 0 && cd >= 0 && ca != cb && cb != cd && ca != cd;
  };
  // search: slot left out of the load quarter (kc), slot riding on the store quarter (mx), lane bits l0 (loads), l1, l2 (stores)
  int best = -1, bkc = 2, bmx = 0;
  std::vector<int> lanes = {cand[0], cand[1], cand[2]};
  const int nc = (int)cand.size();
  if (rd.layout_hint) {                                    // replayed plan: the roles were found when the trace was recorded
    const RoundLayout& h = *rd.layout_hint;
    lanes = {h.lanes[0], h.lanes[1], h.lanes[2]};
    bkc = h.kmap[2]; bmx = h.mmap[0];
    best = 2;
  }
  for (int kc = 0; kc < 3 && best < 2; ++kc) for (int mx = 0; mx < 3 && best < 2; ++mx) {
    const int ka = (kc + 1) % 3, kb = (kc + 2) % 3;
    for (int i0 = 0; i0 < nc && best < 2; ++i0) for (int i1 = 0; i1 < nc && best < 2; ++i1) for (int i2 = i1 + 1; i2 < nc; ++i2) {
      if (i0 == i1 || i0 == i2) continue;
      const int score = (distinct3(rd.slot_pos[ka], rd.slot_pos[kb], cand[i0]) ? 1 : 0) + (distinct3(rd.slot_pos[mx], cand[i1], cand[i2]) ? 1 : 0);
      if (score > best) { best = score; bkc = kc; bmx = mx; lanes = {cand[i0], cand[i1], cand[i2]}; }
      if (best == 2) break;
    }
  }
  {
    int a = (bkc + 1) % 3, b = (bkc + 2) % 3;
    if (a > b) std::swap(a, b);
    rd.kmap[0] = a; rd.kmap[1] = b; rd.kmap[2] = bkc;
    int r0 = (bmx + 1) % 3, r1 = (bmx + 2) % 3;
    if (r0 > r1) std::swap(r0, r1);
    rd.mmap[0] = bmx; rd.mmap[1] = r0; rd.mmap[2] = r1;
  }
  rd.grp_pos = lanes;
  uint64_t lane_mask = 0;
  for (int p : lanes) lane_mask |= 1ULL << p;
  for (int p = 0; p < m; ++p) if (!(((slot_mask | lane_mask | cond) >> p) & 1)) rd.grp_pos.push_back(p);
  for (int p = 0; p < m; ++p) if (((cond >> p) & 1) && !((slot_mask >> p) & 1)) rd.grp_pos.push_back(p);
  const int k = (int)rd.cond_pos.size();
  const size_t nvar = (size_t)1 << k;
  rd.frag.assign(nvar * K3_FRAG_DOUBLES_HOST, 0.0);
  auto pattern_of = [&](int idx, const int (&map)[3]) { return (((idx >> 0) & 1) << map[0]) | (((idx >> 1) & 1) << map[1]) | (((idx >> 2) & 1) << map[2]); };
  int rowp[8], colp[8];                                    // hardware m- / k-index -> slot pattern
  for (int i = 0; i < 8; ++i) { rowp[i] = pattern_of(i, rd.mmap); colp[i] = pattern_of(i, rd.kmap); }
  for (size_t var = 0; var < nvar; ++var) {
    uint64_t fixed = 0;
    for (int j = 0; j < k; ++j) if ((var >> j) & 1) fixed |= 1ULL << rd.cond_pos[j];
    cplx M[8][8];
    for (int row = 0; row < 8; ++row) for (int col = 0; col < 8; ++col) M[row][col] = cplx{row == col ? 1.0 : 0.0, 0.0};
    for (size_t gi = 0; gi < rd.gates.size(); ++gi) {
      const Gate& g = rd.gates[gi];
      // a far phase enters with its far bit = 0; what the bit adds when it is set is in the far table
      small_apply_cols(g, rd.slot_pos, M, 8, far[gi] ? (fixed & ~(gate_bits(g) & not_tile)) : fixed);
    }
    for (int reg = 0; reg < 6; ++reg) {
      double* out = &rd.frag[var * K3_FRAG_DOUBLES_HOST + (size_t)reg * 32];
      for (int lane = 0; lane < 32; ++lane) {
        const cplx z = M[rowp[lane / 4]][colp[lane % 4 + 4 * (reg & 1)]];
        out[lane] = (reg < 2) ? (z.re + z.im) : (reg < 4 ? -z.im : z.re);
      }
    }
  }
  build_far_table(rd.gates, far, rd.slot_pos, m, rd.far, 1);
  build_far_table(rd.gates, far, rd.slot_pos, m, rd.farpre, 2);
  rd.dmma = true;
  rd.k3 = true;
}

// Two rounds in one pass (round kind 3, tile_core.h "paired rounds"): rd.gates on the slot triple S1 = rd.slot_pos, then
// rd.gates2 on the disjoint triple S2, which becomes the batch's three lane bits grp_pos[0..2].  form_rounds guarantees that
// neither block has a condition bit inside the other's slots.  A-fragment registers 0..5 = first block (as build_k3_round),
// 6..11 = second block with its columns in the order the hardware enumerates the lane-group index.
// Bank conflicts: a load quarter varies k-index bits 0, 1 (two S1 bits) and grp_pos[0] (an S2 bit); a store quarter varies
// column bits 1, 2 (two S1 bits, mmap[1], mmap[2]) and m-index bit 0 of the second block (an S2 bit, mmap2[0]).
static void build_k3_pair_round(const Config& cfg, const Stage& st, Round& rd) {
  const int m = st.m;
  uint64_t s1 = 0, s2 = rd.slot_mask2;
  for (int p : rd.slot_pos) s1 |= 1ULL << p;
  std::vector<char> far1, far2;
  classify_far(cfg, rd.gates, s1, m, far1);
  classify_far(cfg, rd.gates2, s2, m, far2);
  const uint64_t cond = round_cond_bits(rd.gates, far1, s1) | round_cond_bits(rd.gates2, far2, s2);
  const uint64_t not_tile = ~((1ULL << m) - 1ULL);
  rd.cond_pos.clear();
  for (int p = 0; p < 64; ++p) if ((cond >> p) & 1) rd.cond_pos.push_back(p);
  uint64_t busy = s1 | s2 | cond;
  for (int p = m - 1; p >= 0 && popc(s1) < 3; --p) if (!((busy >> p) & 1)) { s1 |= 1ULL << p; busy |= 1ULL << p; }
  for (int p = m - 1; p >= 0 && popc(s2) < 3; --p) if (!((busy >> p) & 1)) { s2 |= 1ULL << p; busy |= 1ULL << p; }
  rd.slot_pos.clear();
  std::vector<int> t2;
  for (int p = 0; p < m; ++p) { if ((s1 >> p) & 1) rd.slot_pos.push_back(p); if ((s2 >> p) & 1) t2.push_back(p); }
  const int rb = st.layout_c;
  auto distinct3 = [&](int a, int b, int d) {
    const int ca = chunk_class(a, rb), cb = chunk_class(b, rb), cd = chunk_class(d, rb);
    return ca >= 0 && cb >= 0 && cd >= 0 && ca != cb && cb != cd && ca != cd;
  };
  int bkc = 2, bg0 = 0, bmx = 0, bh = 0;
  const bool hinted = rd.layout_hint != nullptr;
  if (hinted) {
    const RoundLayout& h = *rd.layout_hint;
    bkc = h.kmap[2]; bmx = h.mmap[0]; bh = h.mmap2[0];
    for (int j = 0; j < 3; ++j) if (t2[j] == h.lanes[0]) bg0 = j;
  }
  for (int kc = 0, best = hinted ? 1 : -1; kc < 3 && best < 1; ++kc) for (int g0 = 0; g0 < 3 && best < 1; ++g0) {
    const int sc = distinct3(rd.slot_pos[(kc + 1) % 3], rd.slot_pos[(kc + 2) % 3], t2[g0]) ? 1 : 0;
    if (sc > best) { best = sc; bkc = kc; bg0 = g0; }
  }
  // group bit order of S2: the load-quarter bit first, the other two ascending
  std::vector<int> g2 = {t2[bg0]};
  for (int j = 0; j < 3; ++j) if (j != bg0) g2.push_back(t2[j]);
  for (int mx = 0, best = hinted ? 1 : -1; mx < 3 && best < 1; ++mx) for (int h = 0; h < 3 && best < 1; ++h) {
    const int sc = distinct3(rd.slot_pos[(mx + 1) % 3], rd.slot_pos[(mx + 2) % 3], g2[h]) ? 1 : 0;
    if (sc > best) { best = sc; bmx = mx; bh = h; }
  }
  {
    int a = (bkc + 1) % 3, b = (bkc + 2) % 3;
    if (a > b) std::swap(a, b);
    rd.kmap[0] = a; rd.kmap[1] = b; rd.kmap[2] = bkc;
    int r0 = (bmx + 1) % 3, r1 = (bmx + 2) % 3;
    if (r0 > r1) std::swap(r0, r1);
    rd.mmap[0] = bmx; rd.mmap[1] = r0; rd.mmap[2] = r1;
    int h0 = (bh + 1) % 3, h1 = (bh + 2) % 3;
    if (h0 > h1) std::swap(h0, h1);
    rd.mmap2[0] = bh; rd.mmap2[1] = h0; rd.mmap2[2] = h1;
  }
  rd.grp_pos = g2;
  for (int p = 0; p < m; ++p) if (!(((s1 | s2 | cond) >> p) & 1)) rd.grp_pos.push_back(p);
  for (int p = 0; p < m; ++p) if ((cond >> p) & 1) rd.grp_pos.push_back(p);
  const int k = (int)rd.cond_pos.size();
  const size_t nvar = (size_t)1 << k, FD = 2 * K3_FRAG_DOUBLES_HOST;
  rd.frag.assign(nvar * FD, 0.0);
  auto pattern_of = [&](int idx, const int (&map)[3]) { return (((idx >> 0) & 1) << map[0]) | (((idx >> 1) & 1) << map[1]) | (((idx >> 2) & 1) << map[2]); };
  // A block's matrix depends only on ITS condition bits: of the 2^k variants of the pass only those that differ in them are
  // built, the others copy the fragments of their representative (the variant with the other block's bits cleared).
  const uint64_t condA = round_cond_bits(rd.gates, far1, s1), condB = round_cond_bits(rd.gates2, far2, s2);
  size_t selA = 0, selB = 0;                               // variant-index bits that matter to the first / second block
  for (int j = 0; j < k; ++j) {
    if ((condA >> rd.cond_pos[j]) & 1) selA |= (size_t)1 << j;
    if ((condB >> rd.cond_pos[j]) & 1) selB |= (size_t)1 << j;
  }
  int row1[8], row2[8], col1[8], col2[8];                  // hardware m- / k-index -> pattern of the block
  for (int i = 0; i < 8; ++i) { row1[i] = pattern_of(i, rd.mmap); row2[i] = pattern_of(i, rd.mmap2); col1[i] = pattern_of(i, rd.kmap); col2[i] = 2 * (i & 3) + (i >> 2); }
  auto fill = [&](size_t var, int base_reg, const cplx (*M)[8], const int* rowp, const int* colp) {
    for (int reg = 0; reg < 6; ++reg) {
      double* out = &rd.frag[var * FD + (size_t)(base_reg + reg) * 32];
      for (int lane = 0; lane < 32; ++lane) {
        const cplx z = M[rowp[lane / 4]][colp[lane % 4 + 4 * (reg & 1)]];
        out[lane] = (reg < 2) ? (z.re + z.im) : (reg < 4 ? -z.im : z.re);
      }
    }
  };
  for (size_t var = 0; var < nvar; ++var) {
    uint64_t fixed = 0;
    for (int j = 0; j < k; ++j) if ((var >> j) & 1) fixed |= 1ULL << rd.cond_pos[j];
    const size_t repA = var & selA, repB = var & selB;
    if (repA == var) {
      cplx M1[8][8];
      for (int row = 0; row < 8; ++row) for (int col = 0; col < 8; ++col) M1[row][col] = cplx{row == col ? 1.0 : 0.0, 0.0};
      for (size_t gi = 0; gi < rd.gates.size(); ++gi) {
        const Gate& g = rd.gates[gi];
        small_apply_cols(g, rd.slot_pos, M1, 8, far1[gi] ? (fixed & ~(gate_bits(g) & not_tile)) : fixed);
      }
      fill(var, 0, M1, row1, col1);
    } else {
      std::memcpy(&rd.frag[var * FD], &rd.frag[repA * FD], 6 * 32 * sizeof(double));
    }
    if (repB == var) {
      cplx M2[8][8];
      for (int row = 0; row < 8; ++row) for (int col = 0; col < 8; ++col) M2[row][col] = cplx{row == col ? 1.0 : 0.0, 0.0};
      for (size_t gi = 0; gi < rd.gates2.size(); ++gi) {
        const Gate& g = rd.gates2[gi];
        small_apply_cols(g, g2, M2, 8, far2[gi] ? (fixed & ~(gate_bits(g) & not_tile)) : fixed);
      }
      fill(var, 6, M2, row2, col2);
    } else {
      std::memcpy(&rd.frag[var * FD + 6 * 32], &rd.frag[repB * FD + 6 * 32], 6 * 32 * sizeof(double));
    }
  }
  build_far_table(rd.gates, far1, rd.slot_pos, m, rd.far, 1);
  build_far_table(rd.gates2, far2, g2, m, rd.far2, 1);
  build_far_table(rd.gates, far1, rd.slot_pos, m, rd.farpre, 2);
  build_far_table(rd.gates2, far2, g2, m, rd.farpre2, 2);
  rd.dmma = true;
  rd.k3 = true;
}

static void fuse_round(Round& rd) {
  const int r = (int)rd.slot_pos.size();
  if (r == 0 || rd.gates.size() < 2) return;
  uint64_t slot_mask = 0;
  for (int p : rd.slot_pos) slot_mask |= 1ULL << p;
  const int dim = 1 << r;
  std::vector<Gate> out, run;
  auto flush = [&]() {
    if (run.size() >= 2) {
      Gate d; d.kind = G_DENSE; d.fused = (int)run.size(); d.src_op = run[0].src_op;
      d.dense.assign((size_t)dim * dim, cplx{0, 0});
      for (int col = 0; col < dim; ++col) {           // column `col` = image of basis pattern `col`
        std::vector<cplx> v(dim, cplx{0, 0});
        v[col] = {1, 0};
        for (const Gate& g : run) small_apply(g, rd.slot_pos, v);
        for (int row = 0; row < dim; ++row) d.dense[(size_t)row * dim + col] = v[row];
      }
      out.push_back(std::move(d));
    } else {
      for (Gate& g : run) out.push_back(g);
    }
    run.clear();
  };
  for (Gate& g : rd.gates) {
    if (gate_is_pure(g, slot_mask)) run.push_back(g);
    else { flush(); out.push_back(g); }
  }
  flush();
  rd.gates.swap(out);
}

// Per-gate facts a round candidate needs, computed once per round (not per candidate).
struct RoundGate {
  uint64_t t, d, bits, want;   // non-diagonal targets, diagonal operands, all bits touched, tile-local part of them
  bool can_be_pure, later, reflect;
  bool far_ok;                 // diagonal two-bit gate with exactly one tile-local operand: may ride as a far phase (classify_far)
};

// One candidate round: scan the pending gates in order and take every gate that fits slot bits inside `Rcap` (at most
// MAX_SLOT_BITS of them).  Returns indices into the pending list (taken / rest, order preserved) - no gate is copied.
// Partner search of a paired round (tile_core.h "paired rounds"): touched0 = condition bits of the first block (they count
// against MAX_COND_BITS), avoid = the first block's slot bits - a gate that touches one of them cannot join the second block,
// neither as a target nor as a condition.
static void pick_round(const Config& cfg, const Stage& st, const std::vector<RoundGate>& pg, uint64_t Rcap, std::vector<int>& taken,
                       std::vector<int>& rest, uint64_t& R_out, uint64_t touched0 = 0, uint64_t avoid = 0, bool partner = false) {
  const int rmax = std::min(MAX_SLOT_BITS, st.m);
  const bool use_mma = cfg.dense_mma && st.m >= 6;
  uint64_t R = 0, touched = touched0, bx = 0, bz = 0;   // touched = bits of accepted gates; bx / bz = Blocker state
  uint64_t closed = 0;                                  // slots carrying a far phase: no non-diagonal gate may follow on them
  uint64_t opened = 0;                                  // bits some accepted gate acts on non-diagonally
  taken.clear();
  rest.clear();
  auto fits = [&](uint64_t Rn) { return popc(Rn) <= rmax && (Rn & ~Rcap) == 0; };
  for (size_t gi = 0; gi < pg.size(); ++gi) {
    const RoundGate& g = pg[gi];
    auto block = [&]() { bx |= g.t; bz |= g.d; rest.push_back((int)gi); };
    auto accept = [&](uint64_t Rn) { R = Rn; touched |= g.bits; opened |= g.t; taken.push_back((int)gi); };
    if ((g.t & (bx | bz)) || (g.d & bx)) { block(); continue; }
    if (partner && (g.reflect || (g.bits & avoid))) { block(); continue; }
    if (g.t & closed) { block(); continue; }
    // prefer making the gate *pure* (every tile-local bit it touches becomes a slot bit): pure gates fold into
    // the round's dense block for free; controls / diagonal operands on tile-id or rank bits can never be slots
    if (use_mma && !g.reflect) {
      // tensor-core round: every non-slot bit a gate touches becomes a condition bit (2^k matrix variants)
      auto conds_after = [&](uint64_t Rn) { return popc((touched | g.bits) & ~Rn); };
      if (g.can_be_pure && fits(R | g.want) && conds_after(R | g.want) <= MAX_COND_BITS) { accept(R | g.want); continue; }
      // a diagonal gate that does not fit as pure now: defer it to a later round of this stage if one of its
      // bits will be a slot there anyway (a later gate targets it) - but only a gate that CAN ever be pure (at most rmax
      // tile-local bits; a wider one would be deferred for ever and block everything behind it); otherwise let it ride
      // along as condition bits
      if (g.can_be_pure && g.t == 0 && g.later && popc(g.want) <= rmax) { block(); continue; }
      if (fits(R | g.t) && conds_after(R | g.t) <= MAX_COND_BITS) accept(R | g.t);
      else if (g.far_ok && !(opened & g.want) && fits(R | g.want) && popc((touched | g.want) & ~(R | g.want)) <= MAX_COND_BITS) {
        // the same BEFORE the block (nothing non-diagonal has touched the slot yet in this round): the slot stays open
        R |= g.want; touched |= g.want; taken.push_back((int)gi);
      }
      else if (g.far_ok && fits(R | g.want) && popc((touched | g.want) & ~(R | g.want)) <= MAX_COND_BITS) {
        // out of condition bits: a diagonal two-bit gate with one far operand can still ride as a far phase - its tile-local
        // operand becomes (or is) a slot, the far operand costs nothing; no non-diagonal gate may follow on that slot in this round
        R |= g.want; touched |= g.want; closed |= g.want; taken.push_back((int)gi);
      }
      else if (taken.empty() && Rcap == ~0ULL) accept(R | g.t);     // always make progress (falls back to the interpreter if needed)
      else block();
      continue;
    }
    if (g.can_be_pure && fits(R | g.want)) { accept(R | g.want); continue; }
    if (g.can_be_pure && g.t == 0 && g.later && popc(g.want) <= rmax) { block(); continue; }
    if (fits(R | g.t)) accept(R | g.t);
    else block();
  }
  R_out = R;
}

// Form shared-memory rounds from the gates of one stage (gates already in ext space; targets < m).
// The round budget is counted in quarter rounds: a single round costs 4, a paired pass cfg.pair_cost_q (default 6).
static void materialize_round(const Config& cfg, const Stage& st, Round& rd) {
  if (rd.pair) { build_k3_pair_round(cfg, st, rd); return; }
  if (dmma_eligible(cfg, st, rd)) { if (cfg.mma_form == 0) build_k3_round(cfg, st, rd); else build_dmma_round(cfg, st, rd); }
  else if (cfg.fusion) fuse_round(rd);
}

struct RoundCand { std::vector<int> taken, rest; uint64_t R = 0; };

static void form_rounds(const Config& cfg, Stage& st, std::vector<Gate>& gates, int max_rounds, bool materialize = true) {
  std::vector<Gate> pending = gates;
  const bool search = cfg.fusion && cfg.window_search && cfg.dense_mma && st.m >= 6;
  const bool pairing = search && cfg.pair_rounds && cfg.mma_form == 0 && !cfg.tma && !cfg.direct_store && st.m >= 10;
  const uint64_t tile_mask = (1ULL << st.m) - 1ULL;
  std::vector<RoundGate> pg, pg2;
  std::vector<int> ctaken, crest;
  uint64_t targeted = 0, targeted2 = 0;
  // per-gate facts of the current pending list
  auto prepare = [&]() {
    pg.resize(pending.size());
    uint64_t later_targets = 0;
    targeted = 0;
    for (size_t i = pending.size(); i-- > 0;) {
      const Gate& g = pending[i];
      RoundGate& r = pg[i];
      r.t = g.target_mask(); r.d = g.diag_mask(); r.bits = r.t | r.d; r.want = r.bits & tile_mask;
      r.reflect = g.kind == G_REFLECT;
      r.can_be_pure = cfg.fusion && (r.bits & ~tile_mask) == 0 && g.kind != G_DPOP1 && g.kind != G_REFLECT;
      r.later = (later_targets & r.want) != 0;       // some later gate targets one of its bits
      r.far_ok = cfg.far_phase && cfg.mma_form == 0 && far_kind(g) && popc(r.bits) == 2 && popc(r.want) == 1;
      later_targets |= r.t;
      targeted |= r.t;
    }
  };
  // condition bits a candidate round would have: every non-slot bit its gates touch, far phases excepted (classify_far's rule)
  auto cond_of = [&](const std::vector<RoundGate>& G, const std::vector<int>& tk, uint64_t Rs) {
    uint64_t cond = 0, later = 0, earlier = 0;
    std::vector<char> fl(tk.size(), 0);
    for (size_t k = tk.size(); k-- > 0;) {
      const RoundGate& g = G[tk[k]];
      const uint64_t l = g.bits & Rs;
      if (g.far_ok && popc(l) == 1 && !(later & l)) fl[k] = 1;
      later |= g.t;
    }
    for (size_t k = 0; k < tk.size(); ++k) {
      const RoundGate& g = G[tk[k]];
      const uint64_t l = g.bits & Rs;
      if (!fl[k] && g.far_ok && popc(l) == 1 && !(earlier & l)) fl[k] = 2;
      earlier |= g.t;
      if (!fl[k]) cond |= g.bits & ~Rs;
    }
    return cond;
  };
  // the same facts for a sub-list of the pending gates (what is left once a candidate round has taken its gates)
  auto subset = [&](const std::vector<int>& idx) {
    pg2.resize(idx.size());
    uint64_t later_targets = 0;
    targeted2 = 0;
    for (size_t i = idx.size(); i-- > 0;) {
      RoundGate r = pg[idx[i]];
      r.later = (later_targets & r.want) != 0;
      later_targets |= r.t;
      targeted2 |= r.t;
      pg2[i] = r;
    }
  };
  // The K best rounds of a gate list, best first: greedy (slot bits follow the first gates in line), then - with the search
  // on - every triple of tile-local bits some gate targets; a round is the better the more gates it absorbs (a tensor-core
  // round costs the same however many gates it folds).  partner: slots must avoid `forbid`, see pick_round.
  auto choose = [&](const std::vector<RoundGate>& G, uint64_t targ, bool partner, uint64_t forbid, uint64_t touched0, uint64_t avoid,
                    size_t K, std::vector<RoundCand>& out) {
    out.clear();
    auto offer = [&](std::vector<int>& t, std::vector<int>& r, uint64_t R) {
      if (t.empty()) return;
      for (const RoundCand& c : out) if (c.R == R) return;
      size_t pos = out.size();
      while (pos > 0 && out[pos - 1].taken.size() < t.size()) --pos;      // strictly better moves ahead: ties keep the earlier
      if (pos >= K) return;
      RoundCand c; c.taken = t; c.rest = r; c.R = R;
      out.insert(out.begin() + pos, std::move(c));
      if (out.size() > K) out.pop_back();
    };
    uint64_t cR = 0;
    pick_round(cfg, st, G, partner ? ~forbid : ~0ULL, ctaken, crest, cR, touched0, avoid, partner);
    const size_t greedy_n = ctaken.size();
    offer(ctaken, crest, cR);
    if (search && G.size() > greedy_n) {
      std::vector<int> tb;
      for (int b = 0; b < st.m; ++b) if (((targ & ~forbid) >> b) & 1) tb.push_back(b);
      for (size_t i = 0; i < tb.size(); ++i) for (size_t j = i + 1; j < tb.size(); ++j) for (size_t k = j + 1; k < tb.size(); ++k) {
        const uint64_t cap = (1ULL << tb[i]) | (1ULL << tb[j]) | (1ULL << tb[k]);
        pick_round(cfg, st, G, cap, ctaken, crest, cR, touched0, avoid, partner);
        if (!out.empty() && out.size() >= K && ctaken.size() <= out.back().taken.size()) continue;
        offer(ctaken, crest, cR);
      }
    }
  };
  const int budget_q = 4 * std::max(1, max_rounds);
  int used_q = 4 * (int)st.rounds.size();
  size_t members = st.rounds.size();                 // rounds formed so far, a paired pass counting two
  const size_t K = pairing ? (size_t)std::max(1, cfg.pair_search) : 1;
  const double pair_eff = std::max(100, cfg.pair_eff_pct) / 100.0;   // cost of a paired pass in single rounds (measured: 1.7)
  std::vector<RoundCand> cands, pc;
  while (!pending.empty() && used_q + 4 <= budget_q) {
    prepare();
    choose(pg, targeted, false, 0, 0, 0, K, cands);
    if (cands.empty()) break;                              // nothing fits a round of this stage: leave the rest pending
    // ---- options: the best single round, or one of the K best rounds together with ITS best partner (a round on a disjoint
    // slot triple whose gates touch neither the first round's slots nor - as slots - its condition bits: the two then share
    // one pass over the tile, round kind 3).  The option that absorbs most gates per unit of cost wins.
    size_t best_k = 0;
    bool best_pair = false;
    RoundCand best_b;
    double best_eff = (double)cands[0].taken.size();
    if (pairing && used_q + cfg.pair_cost_q <= budget_q) {
      // the alternative to a pair is two single rounds: the best round and the best round after it, at twice the cost
      if (!cands[0].rest.empty() && used_q + 8 <= budget_q) {
        subset(cands[0].rest);
        choose(pg2, targeted2, false, 0, 0, 0, 1, pc);
        if (!pc.empty()) best_eff = 0.5 * (double)(cands[0].taken.size() + pc[0].taken.size());
      }
      for (size_t k = 0; k < cands.size(); ++k) {
        const RoundCand& a = cands[k];
        if (a.rest.empty()) continue;
        bool ok = true;
        for (int i : a.taken) ok = ok && !pg[i].reflect;
        const uint64_t condA = cond_of(pg, a.taken, a.R);
        const int kl = popc(condA & tile_mask);
        if (!ok || popc(condA) > MAX_COND_BITS || popc(a.R) > 3 || 6 + kl > st.m) continue;
        subset(a.rest);
        choose(pg2, targeted2, true, a.R | (condA & tile_mask), condA, a.R, 1, pc);
        if (pc.empty()) continue;
        const uint64_t condB = cond_of(pg2, pc[0].taken, pc[0].R);
        if (6 + popc((condA | condB) & tile_mask) > st.m || popc(condA | condB) > MAX_COND_BITS) continue;
        const double eff = (double)(a.taken.size() + pc[0].taken.size()) / pair_eff;
        if (eff > best_eff || (!best_pair && eff == best_eff)) { best_eff = eff; best_k = k; best_pair = true; best_b = pc[0]; }
      }
    }
    const RoundCand& A = cands[best_k];
    // a thin round costs as much as a full one: leave its gates to the next sweep, whose tile search starts afresh
    if (search && cfg.round_yield_pct > 0 && members >= 2 && !st.absorbed.empty() &&
        (best_pair ? best_eff : (double)A.taken.size()) * 100 * members < (double)cfg.round_yield_pct * st.absorbed.size()) break;
    Round rd;
    for (int i : A.taken) { rd.gates.push_back(pending[i]); st.absorbed.push_back(pending[i].uid); rd.uids.push_back(pending[i].uid); }
    rd.slot_mask = A.R;
    // unfused mode keeps exactly one gate per round anyway (one gate per stage)
    for (int b = 0; b < st.m; ++b) if ((A.R >> b) & 1) rd.slot_pos.push_back(b);
    std::vector<int> left = A.rest;
    used_q += 4; ++members;
    if (best_pair) {
      rd.pair = true;
      rd.slot_mask2 = best_b.R;
      for (int j : best_b.taken) { const Gate& g = pending[A.rest[j]]; rd.gates2.push_back(g); st.absorbed.push_back(g.uid); rd.uids2.push_back(g.uid); }
      left.clear();
      for (int j : best_b.rest) left.push_back(A.rest[j]);
      used_q += cfg.pair_cost_q - 4; ++members;
    }
    {
      std::vector<Gate> next;
      next.reserve(left.size());
      for (int i : left) next.push_back(std::move(pending[i]));
      pending.swap(next);
    }
    if (materialize) materialize_round(cfg, st, rd);
    st.rounds.push_back(std::move(rd));
  }
}

// Rounds of one stage rebuilt from a recorded trace: same gates per round, same slot bits, fresh matrices.
static void replay_rounds(const Config& cfg, Stage& st, const std::vector<Gate>& gates, const StageTrace& tr) {
  size_t pass = 0;
  for (size_t r = 0; r < tr.round_uids.size(); ++r, ++pass) {
    Round rd;
    if (pass < tr.layouts.size() && tr.layouts[pass].lanes[0] >= 0) rd.layout_hint = &tr.layouts[pass];
    rd.uids = tr.round_uids[r];
    rd.slot_mask = tr.round_slots[r];
    for (int u : rd.uids) {
      for (const Gate& g : gates) if (g.uid == u) { rd.gates.push_back(g); break; }
      st.absorbed.push_back(u);
    }
    for (int b = 0; b < st.m; ++b) if ((rd.slot_mask >> b) & 1) rd.slot_pos.push_back(b);
    if (r < tr.round_pair.size() && tr.round_pair[r] && r + 1 < tr.round_uids.size()) {
      ++r;                                                   // the next recorded round is this one's partner
      rd.pair = true;
      rd.uids2 = tr.round_uids[r];
      rd.slot_mask2 = tr.round_slots[r];
      for (int u : rd.uids2) {
        for (const Gate& g : gates) if (g.uid == u) { rd.gates2.push_back(g); break; }
        st.absorbed.push_back(u);
      }
    }
    materialize_round(cfg, st, rd);
    rd.layout_hint = nullptr;
    st.rounds.push_back(std::move(rd));
  }
}

void plan_structure_key(const Config& cfg, const std::vector<Gate>& gates, const std::vector<int>& perm_in, std::vector<uint64_t>& key,
                        uint64_t support_in) {
  key.clear();
  key.reserve(16 + perm_in.size() + 6 * gates.size());
  const int c[] = {cfg.n_total, cfg.n_local, cfg.rank, cfg.world, cfg.tile_bits, cfg.low_bits, cfg.fusion, cfg.max_stage_cost,
                   cfg.max_stage_rounds, cfg.dense_mma + 16 * cfg.mma_form + 32 * cfg.direct_store, cfg.round_yield_pct, cfg.window_search, cfg.tma, cfg.thin_defer,
                   cfg.pair_rounds + 2 * cfg.pair_eff_pct + 2048 * cfg.pair_cost_q + 65536 * cfg.pair_search + 1048576 * cfg.plan_portfolio + 2097152 * cfg.far_phase};
  for (int v : c) key.push_back((uint64_t)(int64_t)v);
  key.push_back(support_in);
  key.push_back(perm_in.size());
  for (int v : perm_in) key.push_back((uint64_t)v);
  key.push_back(gates.size());
  for (const Gate& g : gates) {
    const uint64_t sign_flip = g.kind == G_DMASK && g.m[0].re == -1.0 && g.m[0].im == 0.0;
    key.push_back((uint64_t)g.kind | ((uint64_t)(uint8_t)(g.t0 + 1) << 8) | ((uint64_t)(uint8_t)(g.t1 + 1) << 16) |
                  ((uint64_t)gate_cost(g) << 24) | (sign_flip << 40));
    key.push_back(g.cmask);
    key.push_back(g.dmask);
    key.push_back(g.dval);
  }
}

// dry: decisions only - no matrices, no encoding, no sink (what the plan portfolio scores its candidates with)
static int schedule_impl(Plan& plan, const std::vector<int>& perm_in, StageSink* sink, PlanTrace* record, const PlanTrace* replay, bool dry) {
  const Config& cfg = plan.cfg;
  const int n = cfg.n_total, nl = cfg.n_local, m = std::min(cfg.tile_bits, nl), L = std::min(cfg.low_bits, m);
  std::vector<int> perm(n);
  for (int b = 0; b < n; ++b) perm[b] = perm_in.empty() ? b : perm_in[b];
  // Budget of one fused sweep.  At 30 qubits a sweep costs ~2.5 ms of HBM / pipeline time plus ~2.1 ms per tensor-core
  // round (profiles/r1e_round_budgets.log).  Long stages end in thin rounds (3-4 gates) that cost as much as full ones,
  // so a stage stops after 5 rounds, or earlier when the next round would hold less than half the stage's average
  // (round_yield_pct); the next sweep's tile search then starts afresh on the leftovers: 72 rounds in 16 sweeps
  // instead of 88 in 10 for the benchmark circuit.
  const int max_cost = cfg.max_stage_cost > 0 ? cfg.max_stage_cost : 800;
  // With paired rounds (two rounds per pass over the tile at ~1.7 x the cost of one) the budget is 7 rounds' worth.
  const bool pairs_on = cfg.fusion && cfg.window_search && cfg.dense_mma && cfg.pair_rounds && cfg.mma_form == 0 && !cfg.tma && !cfg.direct_store;
  const int max_rounds = cfg.max_stage_rounds > 0 ? cfg.max_stage_rounds : (pairs_on ? 7 : 5);
  const double sweep_bytes = 32.0 * std::ldexp(1.0, nl);
  const uint64_t local_mask = (nl >= 64) ? ~0ULL : ((1ULL << nl) - 1);
  const uint64_t tileid_mask = ((nl - m) >= 64) ? ~0ULL : ((1ULL << (nl - m)) - 1);
  uint64_t support = plan.support_in;                    // logical bits that can be 1 where the state is non-zero (plan.h)

  // logical bit space -> physical bit space under the current permutation
  auto to_phys = [&](const Gate& g) {
    Gate p = g;
    auto mp = [&](uint64_t mask) { uint64_t r = 0; for (int b = 0; b < n; ++b) if ((mask >> b) & 1) r |= 1ULL << perm[b]; return r; };
    if (g.t0 >= 0) p.t0 = perm[g.t0];
    if (g.t1 >= 0) p.t1 = perm[g.t1];
    p.cmask = mp(g.cmask);
    if (g.kind == G_DMASK || g.kind == G_DPOP1) { p.dmask = mp(g.dmask); p.dval = mp(g.dval); }
    return p;
  };

  // the program is encoded stage by stage as the plan grows (header word [1] = number of stages, patched at the end)
  plan.words.clear();
  plan.stage_offsets.clear();
  if (replay && replay->words_hint) plan.words.reserve(replay->words_hint + 1024);
  plan.words.push_back(0x51434232ULL);                 // magic "QCB2"
  plan.words.push_back(0);
  plan.words.push_back((uint64_t)n);
  plan.words.push_back((uint64_t)nl);
  plan.n_rounds = 0;
  int sink_rc = QCB_OK;
  auto emit_new_stages = [&]() {
    if (dry) return;
    while (plan.stage_offsets.size() < plan.stages.size() && sink_rc == QCB_OK) {
      const size_t si = plan.stage_offsets.size();
      Stage& s = plan.stages[si];
      plan.stage_offsets.push_back(plan.words.size());
      plan.words.push_back((uint64_t)s.kind);
      if (s.kind == S_TILE) { plan.words.push_back(0); encode_stage(cfg, s, plan.words); for (const Round& r : s.rounds) plan.n_rounds += r.dense_rounds(); }
      else if (s.kind == S_EXCHANGE) { plan.words.push_back((uint64_t)s.gbit | ((uint64_t)s.lbit << 8)); }
      else if (s.kind == S_GROVER) {
        plan.words.push_back((uint64_t)s.marked.size() | ((uint64_t)s.needs_sum << 8));
        for (uint64_t mk : s.marked) plan.words.push_back(mk);
      }
      else { plan.words.push_back(0); }
      plan.words[1] = (uint64_t)plan.stage_offsets.size();
      if (sink) sink_rc = sink->on_stage(plan, si);
    }
  };

  std::vector<int> pending(plan.gates.size());
  for (size_t i = 0; i < pending.size(); ++i) pending[i] = (int)i;
  for (const Gate& g : plan.gates) plan.unfused_bytes += 32.0 * std::ldexp(1.0, n) * g.frac / cfg.world;

  // Build one fused tile stage from the head of `pending`.  `lead` (optional) is an op that must run
  // first on every amplitude (the affine pass of a Grover diffusion).
  // Per-gate facts the stage selection needs (physical bit space), computed once per stage for all tile candidates.
  struct StageGate { uint64_t t, d; int cost; bool reflect, soft_ok, sign_flip; };
  std::vector<StageGate> sg;
  auto prepare_stage_gates = [&](size_t window) {
    const size_t cnt = std::min(pending.size(), window);
    sg.resize(cnt);
    for (size_t i = 0; i < cnt; ++i) {
      const Gate& lg = plan.gates[pending[i]];
      StageGate& r = sg[i];
      r.reflect = lg.kind == G_REFLECT;
      if (r.reflect) { r.t = r.d = 0; r.cost = 0; r.soft_ok = r.sign_flip = false; continue; }
      const Gate g = to_phys(lg);
      r.t = g.target_mask(); r.d = g.diag_mask(); r.cost = gate_cost(g);
      r.soft_ok = (g.kind == G_DMASK || g.kind == G_DTAB1) && popc(r.d) <= 2;
      r.sign_flip = g.kind == G_DMASK && g.m[0].re == -1.0 && g.m[0].im == 0.0;
    }
  };

  // Select the gates of one sweep from the head of `pending`.  A_init = tile bits fixed in advance; with `fixed` the tile
  // may not grow (every target must already be a tile bit), otherwise bits are added greedily in gate order.
  auto select_gates = [&](const Gate* lead, uint64_t A_init, bool fixed, std::vector<int>& taken, uint64_t& A_out) {
    uint64_t A = A_init, bx = 0, bz = 0;                   // bx / bz: bits used non-diagonally / diagonally by skipped gates
    int cost = lead ? 4 : 0;
    taken.clear();
    for (size_t i = 0; i < sg.size(); ++i) {
      const StageGate& g = sg[i];
      if (g.reflect) break;                                  // barrier
      auto block = [&]() { bx |= g.t; bz |= g.d; };
      if ((g.t & ~local_mask) || (g.t & (bx | bz)) || (g.d & bx)) {
        block();
        if (popc(bx | bz) >= n) break;
        continue;
      }
      uint64_t need = g.t & ~A;
      if (fixed && need) { block(); continue; }
      // diagonal gates do not need tile bits, but when the tile has room their operand bits are taken in so
      // that the round fuser can fold them into a dense block (otherwise they ride along as condition bits)
      if (!fixed && cfg.fusion && g.soft_ok) {
        const uint64_t soft = g.d & local_mask & ~A & ~need;
        if (popc(A) + popc(need) + popc(soft) <= m - 1) need |= soft;
        else if (soft && !taken.empty() && !g.sign_flip) {
          // no room: defer a general phase on a bit outside the tile to the sweep that owns the bit (sign flips stay)
          block();
          continue;
        }
      }
      if (popc(A) + popc(need) <= m && ((taken.empty() && !lead) || cost + g.cost <= max_cost)) {
        A |= need; cost += g.cost; taken.push_back((int)i);
      } else {
        block();
      }
    }
    A_out = A;
  };

  // Turn a tile choice (A = tile bits so far, taken = indices into `pending`) into a stage: complete the tile, translate the
  // gates into its ext space and form the rounds.  `absorbed_out` = Gate::uid of the gates the formed rounds hold (the round
  // budget / thin-round cut may leave some of `taken` pending).  materialize = false skips the matrix building (scoring).
  auto make_stage = [&](const Gate* lead, uint64_t A, const std::vector<int>& taken, bool materialize, Stage& st,
                        std::vector<int>& absorbed_out, const StageTrace* tr) {     // taken = gate uids
    st = Stage(); st.kind = S_TILE; st.m = m; st.L = L;
    // single-gate stage: keep its condition bits OUT of the tile so that whole tiles can be skipped
    uint64_t avoid = 0;
    if (taken.size() == 1 && !lead) {
      Gate g = to_phys(plan.gates[taken[0]]);
      if (g.kind == G_MAT1 || g.kind == G_SWAPP || g.kind == G_MAT2 || g.kind == G_DTAB1) avoid = g.cmask;
      else if (g.kind == G_DMASK) avoid = g.dmask;
      avoid &= local_mask;
    }
    for (int b = 0; b < nl && popc(A) < m; ++b) if (!((avoid >> b) & 1)) A |= 1ULL << b;
    for (int b = 0; b < nl && popc(A) < m; ++b) A |= 1ULL << b;
    for (int b = 0; b < nl; ++b) if ((A >> b) & 1) st.tile_pos.push_back(b);
    // TMA mover: one tensor copy moves a run of 2^layout_c amplitudes = the contiguous low tile bits (box rows <= 256
    // => at most 11; at least one 128-byte row => at least 3, else the stage falls back to the LSU mover and layout 0)
    st.layout_c = 0;
    if (cfg.tma) {
      while (st.layout_c < m && st.layout_c < 11 && ((A >> st.layout_c) & 1)) ++st.layout_c;
      if (st.layout_c < 3) st.layout_c = 0;
    }
    std::vector<int> ext_of_phys(64, 0);
    {
      int ti = 0, ni = 0;
      for (int b = 0; b < nl; ++b) { if ((A >> b) & 1) ext_of_phys[b] = ti++; else ext_of_phys[b] = m + ni++; }
      for (int b = nl; b < 64; ++b) ext_of_phys[b] = b;
    }
    // form the rounds; the round budget ends the stage early (gates of the rounds that were not formed stay pending:
    // a prefix of rounds is a valid partial execution because a round only overtakes gates it commutes with)
    std::vector<Gate> eg;
    for (size_t k = 0; k < taken.size(); ++k) {
      eg.push_back(to_ext(to_phys(plan.gates[taken[k]]), ext_of_phys));
      eg.back().uid = taken[k];
    }
    if (lead) { Round r0; r0.gates.push_back(*lead); st.rounds.push_back(r0); }
    if (tr) replay_rounds(cfg, st, eg, *tr);
    else form_rounds(cfg, st, eg, max_rounds, materialize);
    absorbed_out.swap(st.absorbed);
    st.absorbed.clear();
    // Direct store: when the sweep ends in a three-product tensor-core round (and the tile has the default geometry the
    // specialised mover handles), that round writes its results to global memory from registers.  Its store lanes no longer
    // touch shared memory, so the two group bits that ride on the lane index for stores (grp_pos[1], grp_pos[2]) are
    // re-chosen for coalescing instead of bank conflicts: the lowest free tile bits, so that the four lanes of a quad write
    // 64 contiguous bytes.  (grp_pos[0] stays: it belongs to the loads, which still come from the shared tile.)
    st.flags &= ~FLAG_DIRECT_STORE;
    if (materialize && cfg.direct_store && !cfg.tma && m == 12 && L == 4 && !st.rounds.empty() && st.rounds.back().k3 && !st.rounds.back().pair) {
      bool all_mma = true;
      for (const Round& r : st.rounds) all_mma = all_mma && r.dmma;
      if (all_mma) {
        Round& rd = st.rounds.back();
        uint64_t cond = 0;
        for (int p : rd.cond_pos) cond |= 1ULL << p;
        std::vector<int> freeb;                                   // group bits that are not local condition bits, except grp_pos[0]
        for (size_t i = 1; i < rd.grp_pos.size(); ++i) if (!((cond >> rd.grp_pos[i]) & 1)) freeb.push_back(rd.grp_pos[i]);
        if (freeb.size() >= 2) {
          std::sort(freeb.begin(), freeb.end());
          std::vector<int> g;
          g.push_back(rd.grp_pos[0]);
          for (int p : freeb) g.push_back(p);                     // ascending: the two lowest become the store lane bits
          for (size_t i = 1; i < rd.grp_pos.size(); ++i) if ((cond >> rd.grp_pos[i]) & 1) g.push_back(rd.grp_pos[i]);
          rd.grp_pos = g;
          st.flags |= FLAG_DIRECT_STORE;
        }
      }
    }
    st.skip_mask = st.skip_val = 0; st.sweep_fraction = 1.0;
    if (eg.size() == 1 && absorbed_out.size() == 1 && !lead) {
      const Gate& e = eg[0];
      uint64_t cm = 0, cv = 0;
      if (e.kind == G_MAT1 || e.kind == G_SWAPP || e.kind == G_MAT2 || e.kind == G_DTAB1) { cm = e.cmask; cv = e.cmask; }
      else if (e.kind == G_DMASK) { cm = e.dmask; cv = e.dval; }
      st.skip_mask = cm >> m; st.skip_val = cv >> m;
      st.sweep_fraction = std::ldexp(1.0, -popc(st.skip_mask & tileid_mask));
    }
    if (support != ~0ULL && !lead) {
      // known support: an amplitude is zero unless every bit outside the support is 0, so only the tiles whose id has 0 in
      // those bits hold anything (a bit an existing condition already fixes keeps that condition)
      uint64_t zext = 0;
      for (int lb = 0; lb < n; ++lb) {
        if ((support >> lb) & 1) continue;
        const int pb = perm[lb];
        if (pb < nl && ((A >> pb) & 1)) continue;            // a tile bit
        zext |= 1ULL << ext_of_phys[pb];
      }
      zext &= ~(st.skip_mask << m);
      st.skip_mask |= zext >> m;
      st.sweep_fraction = std::ldexp(1.0, -popc(st.skip_mask & tileid_mask));
    }
  };
  // the support after a tile stage: its gates' non-diagonal targets join (the affine pass of a diffusion fills everything)
  auto grow_support = [&](bool lead, const std::vector<int>& uids) {
    if (support == ~0ULL) return;
    if (lead) { support = ~0ULL; return; }
    for (int u : uids) if (u >= 0) support |= plan.gates[u].target_mask();
  };

  auto build_tile_stage = [&](const Gate* lead) -> size_t {
    uint64_t low = 0; for (int k = 0; k < L; ++k) low |= 1ULL << k;
    std::vector<int> taken;
    uint64_t A = low;
    prepare_stage_gates(cfg.fusion ? 4096 : (lead ? 0 : 1));
    select_gates(lead, low, false, taken, A);                // greedy: the tile follows the first gates in line
    if (cfg.fusion && cfg.window_search && !lead && nl > m) {
      // candidate tiles = every contiguous window of m - L physical bits above the low bits (nearest-neighbour circuits
      // leave seams between greedy tiles whose gates then straggle in low-yield sweeps); keep whichever could absorb most
      // gates.  (Scoring the candidates by what their first rounds really hold per estimated millisecond was tried and is
      // no better: 18 sweeps / 72 rounds instead of 16 / 72 for the benchmark circuit.)
      std::vector<int> cand_taken;
      for (int p = L; p + (m - L) <= nl; ++p) {
        uint64_t Aw = low, Aout = 0;
        for (int b = p; b < p + (m - L); ++b) Aw |= 1ULL << b;
        select_gates(lead, Aw, true, cand_taken, Aout);
        if (cand_taken.size() > taken.size()) { taken = cand_taken; A = Aw; }
      }
    }
    if (taken.empty() && !lead) return 0;          // nothing executable in the current layout (multi-GPU: exchange first)
    Stage st; std::vector<int> abs_uids, taken_uids(taken.size());
    for (size_t k = 0; k < taken.size(); ++k) taken_uids[k] = pending[taken[k]];
    make_stage(lead, A, taken_uids, !dry, st, abs_uids, nullptr);
    std::vector<char> absorbed(plan.gates.size(), 0);
    size_t keep = 0;
    for (int u : abs_uids) if (u >= 0 && !absorbed[u]) { absorbed[u] = 1; ++keep; }
    for (size_t k = 0; k < taken.size(); ++k) if (absorbed[pending[taken[k]]]) st.src_gates.push_back(pending[taken[k]]);
    if (record) {
      StageTrace tr; tr.kind = S_TILE; tr.lead = lead != nullptr; tr.tile_bits = A; tr.taken = taken_uids;
      for (size_t r = lead ? 1 : 0; r < st.rounds.size(); ++r) {
        const Round& rd = st.rounds[r];
        tr.round_uids.push_back(rd.uids); tr.round_slots.push_back(rd.slot_mask); tr.round_pair.push_back(rd.pair ? 1 : 0);
        if (rd.pair) { tr.round_uids.push_back(rd.uids2); tr.round_slots.push_back(rd.slot_mask2); tr.round_pair.push_back(0); }
        RoundLayout lay;
        for (int j = 0; j < 3; ++j) { lay.kmap[j] = rd.kmap[j]; lay.mmap[j] = rd.mmap[j]; lay.mmap2[j] = rd.mmap2[j]; lay.lanes[j] = -1; }
        if (rd.k3 && rd.grp_pos.size() >= 3) for (int j = 0; j < 3; ++j) lay.lanes[j] = rd.grp_pos[j];
        tr.layouts.push_back(lay);
      }
      record->stages.push_back(std::move(tr));
    }
    plan.stages.push_back(st);
    plan.algorithmic_bytes += sweep_bytes * st.sweep_fraction;
    grow_support(lead != nullptr, abs_uids);
    std::vector<int> rest;
    for (size_t i = 0; i < pending.size(); ++i) if (!absorbed[pending[i]]) rest.push_back(pending[i]);
    pending.swap(rest);
    return keep + (lead ? 1 : 0);
  };

  // phase oracle = sign flip of exactly one basis state (all n bits compared)
  const uint64_t full_mask = (n >= 64) ? ~0ULL : ((1ULL << n) - 1);
  auto is_oracle_flip = [&](const Gate& g) {
    return g.kind == G_DMASK && g.dmask == full_mask && g.m[0].re == -1.0 && g.m[0].im == 0.0;
  };
  // uids = {diffusion, oracle...}: the marked indices in the current physical layout, kept when they live on this rank
  auto make_grover_stage = [&](const std::vector<int>& uids, bool needs_sum) {
    Stage s; s.kind = S_GROVER; s.needs_sum = needs_sum; s.src_gates = uids;
    for (size_t i = 1; i < uids.size(); ++i) {
      const uint64_t idx = to_phys(plan.gates[uids[i]]).dval;
      if (nl >= 64 || (idx >> nl) == (uint64_t)cfg.rank) s.marked.push_back(idx & local_mask);
    }
    return s;
  };
  if (record) { record->stages.clear(); plan_structure_key(cfg, plan.gates, perm_in, record->key, plan.support_in); }
  if (replay) {
    // ---- the decisions come from a trace of a structurally identical circuit: no searching
    Gate lead_gate;
    for (const StageTrace& tr : replay->stages) {
      emit_new_stages();
      if (sink_rc != QCB_OK) { plan.error = "stage sink failed"; return sink_rc; }
      if (tr.kind == S_SUM) {
        Stage s; s.kind = S_SUM; s.src_gates.push_back(tr.lead_uid);
        plan.stages.push_back(s);
        plan.algorithmic_bytes += 0.5 * sweep_bytes;
      } else if (tr.kind == S_GROVER) {
        plan.stages.push_back(make_grover_stage(tr.taken, tr.needs_sum));
        plan.algorithmic_bytes += sweep_bytes;
        support = ~0ULL;
      } else if (tr.kind == S_EXCHANGE) {
        std::vector<int> logical_of(n);
        for (int b = 0; b < n; ++b) logical_of[perm[b]] = b;
        Stage s; s.kind = S_EXCHANGE; s.gbit = tr.gbit; s.lbit = tr.lbit;
        plan.stages.push_back(s);
        plan.n_exchanges++;
        std::swap(perm[logical_of[tr.gbit]], perm[logical_of[tr.lbit]]);
      } else {
        const Gate* lead = nullptr;
        if (tr.lead) { lead_gate = Gate(); lead_gate.kind = G_REFLECT; lead_gate.src_op = plan.gates[tr.lead_uid].src_op; lead = &lead_gate; }
        Stage st; std::vector<int> abs_uids;
        make_stage(lead, tr.tile_bits, tr.taken, true, st, abs_uids, &tr);
        std::vector<char> absorbed(plan.gates.size(), 0);
        for (int u : abs_uids) if (u >= 0) absorbed[u] = 1;
        for (int u : tr.taken) if (absorbed[u]) st.src_gates.push_back(u);
        plan.stages.push_back(st);
        plan.algorithmic_bytes += sweep_bytes * st.sweep_fraction;
        grow_support(lead != nullptr, abs_uids);
      }
    }
    pending.clear();
  }
  while (!pending.empty()) {
    emit_new_stages();
    if (sink_rc != QCB_OK) { plan.error = "stage sink failed"; return sink_rc; }
    // ---- Grover diffusion 2|s><s| - I: a read-only sum sweep, then a' = 2*mean - a opens the next sweep
    if (plan.gates[pending[0]].kind == G_REFLECT) {
      // the sum is already on the device when the previous stage was a fused Grover pass that computed it
      const bool have_sum = !plan.stages.empty() && plan.stages.back().kind == S_GROVER && plan.stages.back().needs_sum;
      // diffusion followed by nothing but phase oracles up to the next diffusion (or the end): one streaming pass
      size_t k = 1;
      while (k < pending.size() && k <= (size_t)MAX_GROVER_MARKED && is_oracle_flip(plan.gates[pending[k]])) ++k;
      const bool next_reflect = k < pending.size() && plan.gates[pending[k]].kind == G_REFLECT;
      if (cfg.fusion && (next_reflect || k == pending.size())) {
        if (!have_sum) {
          Stage s; s.kind = S_SUM; s.src_gates.push_back(pending[0]);
          plan.stages.push_back(s);
          plan.algorithmic_bytes += 0.5 * sweep_bytes;
          if (record) { StageTrace tr; tr.kind = S_SUM; tr.lead_uid = pending[0]; record->stages.push_back(tr); }
        }
        std::vector<int> uids(pending.begin(), pending.begin() + k);
        plan.stages.push_back(make_grover_stage(uids, next_reflect));
        plan.algorithmic_bytes += sweep_bytes;
        support = ~0ULL;
        if (record) { StageTrace tr; tr.kind = S_GROVER; tr.taken = uids; tr.needs_sum = next_reflect; record->stages.push_back(tr); }
        pending.erase(pending.begin(), pending.begin() + k);
        continue;
      }
      if (have_sum) {
        Gate a; a.kind = G_REFLECT; a.src_op = plan.gates[pending[0]].src_op;
        const int reflect_uid = pending[0];
        pending.erase(pending.begin());
        build_tile_stage(&a);
        if (record) record->stages.back().lead_uid = reflect_uid;
        continue;
      }
      Stage s; s.kind = S_SUM; s.src_gates.push_back(pending[0]);
      plan.stages.push_back(s);
      plan.algorithmic_bytes += 0.5 * sweep_bytes;
      Gate a; a.kind = G_REFLECT; a.src_op = plan.gates[pending[0]].src_op;
      const int reflect_uid = pending[0];
      if (record) { StageTrace tr; tr.kind = S_SUM; tr.lead_uid = reflect_uid; record->stages.push_back(tr); }
      pending.erase(pending.begin());
      build_tile_stage(&a);
      if (record) record->stages.back().lead_uid = reflect_uid;
      continue;
    }
    // ---- everything executable in the current layout goes first: gates with a non-diagonal target on a global
    // physical bit (and whatever depends on them) are skipped by the stage builder, so an exchange is only paid for
    // when no gate at all can run without it
    {
      // Do not run a THIN stage (fewer gates than thin_defer, default 12) while gates wait for an exchange: its gates ride
      // along in the fuller sweeps after the exchange.  Round 1 kept this off because every extra exchange cost a pairwise
      // half-slice transfer; with up to three global qubits moving in ONE in-place pass a moderate threshold pays: the
      // 33-qubit benchmark on 8 GPUs goes from 23 sweeps / 89 rounds to 21 / 86 with the same single exchange pass, QFT-26 on
      // 8 GPUs from 23 sweeps / 5 exchange passes to 19 / 6.  Larger thresholds trade sweeps for many more exchanges
      // (QFT-26: 10 passes at 16, 18 at 24) and are not worth it.
      const bool may_defer = cfg.world > 1 && cfg.fusion && cfg.thin_defer > 0 &&
                             !(plan.stages.size() && plan.stages.back().kind == S_EXCHANGE);
      std::vector<int> saved_pending;
      double saved_bytes = plan.algorithmic_bytes;
      const size_t saved_traces = record ? record->stages.size() : 0;
      if (may_defer) saved_pending = pending;
      const uint64_t saved_support = support;
      const size_t got = build_tile_stage(nullptr);
      if (got && may_defer && (int)got < cfg.thin_defer && got < pending.size() + got) {
        // is anything blocked on a global bit?
        bool blocked = false;
        for (size_t i = 0; i < pending.size() && i < 4096 && !blocked; ++i)
          blocked = (to_phys(plan.gates[pending[i]]).target_mask() & ~local_mask) != 0;
        if (blocked) {
          plan.stages.pop_back();
          plan.algorithmic_bytes = saved_bytes;
          support = saved_support;
          if (record) record->stages.resize(saved_traces);
          pending.swap(saved_pending);
        } else continue;
      } else if (got) continue;
    }
    // ---- multi-GPU remap: the head-of-line gate targets a global bit.  Swap it with the local bit (bit 12 and up: rows of
    // >= 64 KiB for the strided copies) whose logical occupant is needed latest as a non-diagonal target.  Every other
    // global bit that pending gates target as well is swapped in the same breath when a partner exists that no pending
    // gate targets any more: those exchanges have to happen anyway, and done back to back they spare the thin sweeps that
    // would otherwise run between them (33-qubit benchmark on 8 GPUs: 6 exchanges + 27 sweeps -> 3 + 21).
    {
      Gate g0; uint64_t gt = 0;
      for (size_t i = 0; i < pending.size() && !gt; ++i) { g0 = to_phys(plan.gates[pending[i]]); gt = g0.target_mask() & ~local_mask; }
      size_t trailing_exchanges = 0;
      for (size_t i = plan.stages.size(); i-- > 0 && plan.stages[i].kind == S_EXCHANGE;) ++trailing_exchanges;
      if (!gt || trailing_exchanges > (size_t)(2 * n)) { plan.error = "scheduler made no progress"; return QCB_ERR_INVALID; }
      const int lo_cand = std::max(L, std::min(nl - 8, 12));
      const size_t window = std::min<size_t>(pending.size(), 4096);
      const size_t never = pending.size() + 1;
      std::vector<int> logical_of(n);
      auto refresh = [&]() { for (int b = 0; b < n; ++b) logical_of[perm[b]] = b; };
      refresh();
      // next use (index into pending) of the logical occupant of physical bit `pb` as a non-diagonal target
      auto next_use = [&](int pb) {
        const int lb = logical_of[pb];
        for (size_t i = 0; i < window; ++i)
          if ((plan.gates[pending[i]].target_mask() >> lb) & 1) return i;
        return never;
      };
      auto emit_exchange = [&](int gbit, int lbit) {
        Stage s; s.kind = S_EXCHANGE; s.gbit = gbit; s.lbit = lbit;
        plan.stages.push_back(s);
        plan.n_exchanges++;
        if (record) { StageTrace tr; tr.kind = S_EXCHANGE; tr.gbit = gbit; tr.lbit = lbit; record->stages.push_back(tr); }
        std::swap(perm[logical_of[gbit]], perm[logical_of[lbit]]);
        refresh();
      };
      const int gbit = 63 - __builtin_clzll(gt);
      int best = -1; size_t best_next = 0;
      for (int lo = lo_cand; best < 0; lo = 0) {             // tiny slices: fall back to any local bit
        for (int cand = nl - 1; cand >= lo && cand >= 0; --cand) {
          if ((g0.target_mask() >> cand) & 1) continue;
          const size_t next = next_use(cand);
          if (best < 0 || next > best_next) { best = cand; best_next = next; }
        }
        if (lo == 0) break;
      }
      if (best < 0) { plan.error = "no local qubit available for remap"; return QCB_ERR_INVALID; }
      emit_exchange(gbit, best);
      for (int gb = n - 1; gb >= nl; --gb) {
        if (next_use(gb) == never) continue;               // nothing pending targets this global bit
        int partner = -1;
        for (int cand = nl - 1; cand >= lo_cand && cand >= 0; --cand)
          if (next_use(cand) == never) { partner = cand; break; }
        if (partner < 0) break;                            // no free partner: leave the choice to the next trigger
        emit_exchange(gb, partner);
      }
    }
  }

  plan.perm_out = perm;
  plan.support_out = support;
  emit_new_stages();
  if (sink_rc != QCB_OK) { plan.error = "stage sink failed"; return sink_rc; }
  if (record) record->words_hint = plan.words.size();
  return QCB_OK;
}

// Estimated device time of a plan in milliseconds at 2^30 amplitudes per GPU, from the cost model fitted to the B200
// measurements of round 2 (profiles/r2l_ab.log: 13 plans of the same circuit, residual <= 2 %): a sweep costs 1.52 ms on top
// of its passes (mover traffic, fill / drain), a single tensor-core round 2.0 ms, a paired pass 3.39 ms; an interpreter round
// is charged like a paired pass, a qubit exchange like 8 ms (2-GPU measurement, 12.4 ms for one qubit, less per qubit when
// several move in one pass).  Only the RATIOS matter: the portfolio compares plans of one circuit on one machine.
static double plan_cost_ms(const Plan& plan) {
  double c = 0.0;
  for (const Stage& st : plan.stages) {
    if (st.kind == S_EXCHANGE) { c += 8.0; continue; }
    if (st.kind == S_SUM) { c += 2.6; continue; }
    if (st.kind == S_GROVER) { c += 5.3; continue; }
    double t = 1.52;
    for (const Round& rd : st.rounds) t += rd.pair ? 3.39 : 2.0;
    // a sweep cannot beat its HBM traffic: 32 B per amplitude at the 0.9 of the copy peak that two-round sweeps reach
    // (bench.py: roofline.hbm_bound_config, 37 sweeps in 216.6 ms) - thin tail sweeps are not as cheap as their rounds
    t = std::max(t, 5.8);
    c += t * st.sweep_fraction;
  }
  return c;
}

// Plan portfolio: the greedy stage / round builders are sensitive to their budgets (how many rounds a sweep may hold, when a
// partner round is worth a paired pass), and which setting wins depends on the circuit (+-4 % on the brickwork circuits of 26
// to 33 qubits).  For large circuits the scheduler therefore makes its decisions under a handful of settings - decisions only,
// no matrices, a few milliseconds each - keeps the plan the cost model likes best and replays that one through the normal path.
int schedule(Plan& plan, const std::vector<int>& perm_in, StageSink* sink, PlanTrace* record, const PlanTrace* replay) {
  const Config base = plan.cfg;
  const bool pairs_on = base.fusion && base.window_search && base.dense_mma && base.pair_rounds && base.mma_form == 0 && !base.tma && !base.direct_store;
  if (replay || !base.plan_portfolio || !pairs_on || base.max_stage_rounds > 0 || base.n_local < 24 || plan.gates.size() < 128)
    return schedule_impl(plan, perm_in, sink, record, replay, false);
  // yield = round_yield_pct (when a stage ends early), 0 = the handle's setting.  The first NBASE entries are the settings every
  // hardware measurement of round 2 was taken with; the others vary the yield threshold, which the cost model prefers by 2 - 7 % on
  // the brickwork circuits of 28, 29, 31, 32 qubits and on 32 / 36 qubits across 4 / 8 GPUs and never at 30 (host-only finding of the
  // last session of the round, no GPU time left to confirm it): a variant is taken only when it is predicted to be at least 2 %
  // cheaper than the best base candidate, the fit's residual.
  struct Knobs { int rounds, cost_q, eff_pct, search, pairs, yield; };
  static const Knobs kn[] = {{7, 7, 170, 1, 1, 0}, {7, 7, 150, 4, 1, 0}, {6, 6, 170, 1, 1, 0}, {8, 7, 160, 4, 1, 0}, {5, 6, 170, 1, 0, 0},
                             {7, 7, 170, 1, 1, 35}, {7, 7, 150, 4, 1, 35}, {6, 6, 170, 1, 1, 35}, {8, 7, 160, 4, 1, 35},
                             {7, 7, 170, 1, 1, 75}, {7, 7, 150, 4, 1, 75}, {6, 6, 170, 1, 1, 75}, {8, 7, 160, 4, 1, 75}};
  constexpr int NBASE = 5;
  auto with = [&](const Knobs& k) {
    Config c = base;
    c.max_stage_rounds = k.rounds; c.pair_cost_q = k.cost_q; c.pair_eff_pct = k.eff_pct; c.pair_search = k.search; c.pair_rounds = k.pairs;
    if (k.yield) c.round_yield_pct = k.yield;
    return c;
  };
  // the candidates are independent: one host thread each (the scheduler keeps no global state)
  static const bool yields_on = !(std::getenv("QCB_PORTFOLIO_YIELDS") && std::atoi(std::getenv("QCB_PORTFOLIO_YIELDS")) == 0);
  const int NK = (base.round_yield_pct == 50 && yields_on) ? (int)(sizeof kn / sizeof kn[0]) : NBASE;   // a pinned yield stays pinned
  constexpr int NKMAX = (int)(sizeof kn / sizeof kn[0]);
  PlanTrace traces[NKMAX];
  double costs[NKMAX];
  int rcs[NKMAX];
  {
    auto work = [&](int i) {
      Plan t;
      t.cfg = with(kn[i]);
      t.gates = plan.gates;
      t.support_in = plan.support_in;
      rcs[i] = schedule_impl(t, perm_in, nullptr, &traces[i], nullptr, true);
      costs[i] = rcs[i] == QCB_OK ? plan_cost_ms(t) : 0.0;
    };
    std::vector<std::thread> th;
    th.reserve(NK);
    for (int i = 0; i < NK; ++i) {
      try { th.emplace_back(work, i); }
      catch (const std::system_error&) { work(i); }        // no thread to be had (pid limit of a container): score it here
    }
    for (auto& t : th) t.join();
  }
  int best = -1;
  double best_cost = 0.0;
  for (int i = 0; i < NBASE; ++i)
    if (rcs[i] == QCB_OK && (best < 0 || costs[i] < best_cost)) { best = i; best_cost = costs[i]; }
  {
    int var = -1;
    for (int i = NBASE; i < NK; ++i)
      if (rcs[i] == QCB_OK && (var < 0 || costs[i] < costs[var])) var = i;
    if (var >= 0 && (best < 0 || costs[var] < 0.98 * best_cost)) { best = var; best_cost = costs[var]; }
  }
  static const bool debug = std::getenv("QCB_PORTFOLIO_DEBUG") && std::atoi(std::getenv("QCB_PORTFOLIO_DEBUG"));
  if (debug) {
    std::fprintf(stderr, "[portfolio] n_local %d gates %zu:", base.n_local, plan.gates.size());
    for (int i = 0; i < NK; ++i) std::fprintf(stderr, " %s%.2f", i == best ? "*" : "", rcs[i] == QCB_OK ? costs[i] : -1.0);
    std::fprintf(stderr, "\n");
  }
  PlanTrace best_trace;
  if (best >= 0) best_trace = std::move(traces[best]);
  if (best < 0) return schedule_impl(plan, perm_in, sink, record, nullptr, false);
  plan.cfg = with(kn[best]);
  const int rc = schedule_impl(plan, perm_in, sink, nullptr, &best_trace, false);
  plan.cfg = base;
  if (record) {
    *record = std::move(best_trace);
    plan_structure_key(base, plan.gates, perm_in, record->key, plan.support_in);
    record->words_hint = plan.words.size();
    // the candidates were scored without matrices, so their traces carry no lane roles: take them from the plan just built
    size_t si = 0;
    for (StageTrace& tr : record->stages) {
      if (tr.kind != S_TILE) { if (tr.kind == S_EXCHANGE || tr.kind == S_SUM || tr.kind == S_GROVER) ++si; continue; }
      if (si >= plan.stages.size()) break;
      const Stage& st = plan.stages[si++];
      tr.layouts.clear();
      for (size_t r = tr.lead ? 1 : 0; r < st.rounds.size(); ++r) {
        const Round& rd = st.rounds[r];
        RoundLayout lay;
        for (int j = 0; j < 3; ++j) { lay.kmap[j] = rd.kmap[j]; lay.mmap[j] = rd.mmap[j]; lay.mmap2[j] = rd.mmap2[j]; lay.lanes[j] = -1; }
        if (rd.k3 && rd.grp_pos.size() >= 3) for (int j = 0; j < 3; ++j) lay.lanes[j] = rd.grp_pos[j];
        tr.layouts.push_back(lay);
      }
    }
  }
  return rc;
}

}  // namespace qcb
