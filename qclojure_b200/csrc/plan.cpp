// plan.cpp — lowering of the reference gate vocabulary, the fusion scheduler and the program encoder.
// Host-only (compiled by nvcc into libqcb200.so and by g++ into the CPU test emulator).
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>

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
  k.fusion = c.fusion;
  k.strict = c.strict_parity;
  k.max_stage_cost = c.max_stage_cost;
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
      bool perm = g.m[0].re == 0 && g.m[0].im == 0 && g.m[3].re == 0 && g.m[3].im == 0;
      return perm ? 2 : 8;
    }
    case G_MAT2: return 32;
    case G_SWAPP: return 2;
    case G_DMASK: return 2;
    case G_DTAB1: return 4;
    case G_DPOP1: return 3;
    default: return 8;
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

// choose lane positions so that the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
// under the swizzle  phys = i ^ (((i>>3) ^ (i>>6) ^ (i>>9)) & 7)   (tile_core.h: swz)
static void choose_lanes(int m, const std::vector<int>& slots, std::vector<int>& lanes) {
  lanes.clear();
  uint64_t used = 0;
  for (int s : slots) used |= 1ULL << s;
  int want = std::min(3, m - (int)slots.size());
  for (int k = 0; k < 3 && (int)lanes.size() < want; ++k) {
    int pick = -1;
    for (int c = k; c < m; c += 3)
      if (!((used >> c) & 1)) { pick = c; break; }
    if (pick >= 0) { lanes.push_back(pick); used |= 1ULL << pick; }
  }
  for (int c = 0; c < m && (int)lanes.size() < want; ++c)
    if (!((used >> c) & 1)) { lanes.push_back(c); used |= 1ULL << c; }
}

static void put_gate_words(std::vector<uint64_t>& w, const Gate& g, const std::vector<int>& slot_pos) {
  auto slot_of = [&](int pos) {
    for (size_t j = 0; j < slot_pos.size(); ++j) if (slot_pos[j] == pos) return (int)j;
    return -1;
  };
  auto dbl = [](double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; };
  size_t base = w.size();
  int nslots = (g.kind == G_MAT2) ? 3 : 1;
  w.resize(base + (size_t)nslots * OP_WORDS, 0);
  uint64_t kind = 0, j0 = 0, j1 = 0;
  switch (g.kind) {
    case G_MAT1: kind = D_MAT1; j0 = slot_of(g.t0); break;
    case G_MAT2: kind = D_MAT2; j0 = slot_of(g.t0); j1 = slot_of(g.t1); break;
    case G_SWAPP: kind = D_SWAPP; j0 = slot_of(g.t0); j1 = slot_of(g.t1); break;
    case G_DMASK: kind = D_DMASK; break;
    case G_DTAB1: kind = D_DTAB1; break;
    case G_DPOP1: kind = D_DPOP1; break;
    default: kind = D_AFFINE; break;
  }
  w[base + 0] = kind | (j0 << 8) | (j1 << 16) | ((uint64_t)nslots << 24);
  if (g.kind == G_DMASK || g.kind == G_DPOP1) { w[base + 1] = g.dmask; w[base + 2] = g.dval; }
  else if (g.kind == G_DTAB1) { w[base + 1] = g.cmask; w[base + 2] = (uint64_t)g.t0; }
  else { w[base + 1] = g.cmask; w[base + 2] = g.dval; }
  int nc = (g.kind == G_MAT2) ? 16 : 4;
  for (int i = 0; i < nc; ++i) { w[base + 4 + 2 * i] = dbl(g.m[i].re); w[base + 5 + 2 * i] = dbl(g.m[i].im); }
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
  // rounds
  size_t rbase = words.size();
  words.resize(rbase + st.rounds.size() * ROUND_WORDS, 0);
  for (size_t r = 0; r < st.rounds.size(); ++r) {
    Round& rd = st.rounds[r];
    std::vector<int> lanes;
    choose_lanes(m, rd.slot_pos, lanes);
    size_t ob = words.size();
    for (const Gate& g : rd.gates) put_gate_words(words, g, rd.slot_pos);
    size_t rb = rbase + r * ROUND_WORDS;
    words[rb + 0] = rd.slot_pos.size();
    words[rb + 1] = (words.size() - ob) / OP_WORDS;
    words[rb + 2] = ob - base;
    words[rb + 3] = lanes.size();
    for (size_t j = 0; j < rd.slot_pos.size(); ++j) words[rb + 4 + j] = (uint64_t)rd.slot_pos[j];
    for (size_t j = 0; j < lanes.size(); ++j) words[rb + 7 + j] = (uint64_t)lanes[j];
    std::vector<int> ins(rd.slot_pos);
    ins.insert(ins.end(), lanes.begin(), lanes.end());
    std::sort(ins.begin(), ins.end());
    words[rb + 10] = ins.size();
    for (size_t j = 0; j < ins.size(); ++j) words[rb + 11 + j] = (uint64_t)ins[j];
  }
  words[base + 40] = words.size() - base;
}

// Form shared-memory rounds from the gates of one stage (gates already in ext space; targets < m).
static void form_rounds(const Config& cfg, Stage& st, std::vector<Gate>& gates) {
  std::vector<Gate> pending = gates;
  const int rmax = std::min(MAX_SLOT_BITS, st.m);
  while (!pending.empty()) {
    Round rd;
    uint64_t R = 0;
    Blocker bl;
    std::vector<Gate> rest;
    for (const Gate& g : pending) {
      if (bl.conflicts(g)) { bl.block(g); rest.push_back(g); continue; }
      uint64_t t = g.target_mask();
      if (popc(R | t) <= rmax) { R |= t; rd.gates.push_back(g); }
      else { bl.block(g); rest.push_back(g); }
    }
    // unfused mode keeps exactly one gate per round anyway (one gate per stage)
    for (int b = 0; b < st.m; ++b) if ((R >> b) & 1) rd.slot_pos.push_back(b);
    // pad with extra slot bits when the tile is so small that fewer than 8 lanes exist: not needed
    st.rounds.push_back(std::move(rd));
    pending.swap(rest);
  }
  (void)cfg;
}

int schedule(Plan& plan, const std::vector<int>& perm_in) {
  const Config& cfg = plan.cfg;
  const int n = cfg.n_total, nl = cfg.n_local, m = std::min(cfg.tile_bits, nl), L = std::min(cfg.low_bits, m);
  std::vector<int> perm(n);
  for (int b = 0; b < n; ++b) perm[b] = perm_in.empty() ? b : perm_in[b];
  const int max_cost = cfg.max_stage_cost > 0 ? cfg.max_stage_cost : 96;
  const double sweep_bytes = 32.0 * std::ldexp(1.0, nl);
  const uint64_t local_mask = (nl >= 64) ? ~0ULL : ((1ULL << nl) - 1);
  const uint64_t tileid_mask = ((nl - m) >= 64) ? ~0ULL : ((1ULL << (nl - m)) - 1);

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

  std::vector<int> pending(plan.gates.size());
  for (size_t i = 0; i < pending.size(); ++i) pending[i] = (int)i;
  for (const Gate& g : plan.gates) plan.unfused_bytes += 32.0 * std::ldexp(1.0, n) * g.frac / cfg.world;

  // Build one fused tile stage from the head of `pending`.  `lead` (optional) is an op that must run
  // first on every amplitude (the affine pass of a Grover diffusion).
  auto build_tile_stage = [&](const Gate* lead) {
    Stage st; st.kind = S_TILE; st.m = m; st.L = L;
    uint64_t A = 0; for (int k = 0; k < L; ++k) A |= 1ULL << k;
    Blocker bl; int cost = lead ? 4 : 0;
    std::vector<int> taken;
    const size_t window = cfg.fusion ? 4096 : (lead ? 0 : 1);
    for (size_t i = 0; i < pending.size() && i < window; ++i) {
      const Gate& lg = plan.gates[pending[i]];
      if (lg.kind == G_REFLECT) break;                       // barrier
      Gate g = to_phys(lg);
      if ((g.target_mask() & ~local_mask) || bl.conflicts(g)) {
        bl.block(g);
        if (popc(bl.x | bl.z) >= n) break;
        continue;
      }
      uint64_t need = g.target_mask() & ~A;
      int c = gate_cost(g);
      if (popc(A) + popc(need) <= m && ((taken.empty() && !lead) || cost + c <= max_cost)) {
        A |= need; cost += c; taken.push_back((int)i);
      } else {
        bl.block(g);
      }
    }
    // single-gate stage: keep its condition bits OUT of the tile so that whole tiles can be skipped
    uint64_t avoid = 0;
    if (taken.size() == 1 && !lead) {
      Gate g = to_phys(plan.gates[pending[taken[0]]]);
      if (g.kind == G_MAT1 || g.kind == G_SWAPP || g.kind == G_MAT2 || g.kind == G_DTAB1) avoid = g.cmask;
      else if (g.kind == G_DMASK) avoid = g.dmask;
      avoid &= local_mask;
    }
    for (int b = 0; b < nl && popc(A) < m; ++b) if (!((avoid >> b) & 1)) A |= 1ULL << b;
    for (int b = 0; b < nl && popc(A) < m; ++b) A |= 1ULL << b;
    for (int b = 0; b < nl; ++b) if ((A >> b) & 1) st.tile_pos.push_back(b);
    std::vector<int> ext_of_phys(64, 0);
    {
      int ti = 0, ni = 0;
      for (int b = 0; b < nl; ++b) { if ((A >> b) & 1) ext_of_phys[b] = ti++; else ext_of_phys[b] = m + ni++; }
      for (int b = nl; b < 64; ++b) ext_of_phys[b] = b;
    }
    std::vector<Gate> eg;
    std::vector<char> tk(pending.size(), 0);
    for (int i : taken) { tk[i] = 1; eg.push_back(to_ext(to_phys(plan.gates[pending[i]]), ext_of_phys)); st.src_gates.push_back(pending[i]); }
    if (eg.size() == 1 && !lead) {
      const Gate& e = eg[0];
      uint64_t cm = 0, cv = 0;
      if (e.kind == G_MAT1 || e.kind == G_SWAPP || e.kind == G_MAT2 || e.kind == G_DTAB1) { cm = e.cmask; cv = e.cmask; }
      else if (e.kind == G_DMASK) { cm = e.dmask; cv = e.dval; }
      st.skip_mask = cm >> m; st.skip_val = cv >> m;
      st.sweep_fraction = std::ldexp(1.0, -popc(st.skip_mask & tileid_mask));
    }
    if (lead) { Round r0; r0.gates.push_back(*lead); st.rounds.push_back(r0); }
    form_rounds(cfg, st, eg);
    plan.stages.push_back(st);
    plan.algorithmic_bytes += sweep_bytes * st.sweep_fraction;
    std::vector<int> rest;
    for (size_t i = 0; i < pending.size(); ++i) if (!tk[i]) rest.push_back(pending[i]);
    pending.swap(rest);
  };

  while (!pending.empty()) {
    // ---- Grover diffusion 2|s><s| - I: a read-only sum sweep, then a' = 2*mean - a opens the next sweep
    if (plan.gates[pending[0]].kind == G_REFLECT) {
      Stage s; s.kind = S_SUM; s.src_gates.push_back(pending[0]);
      plan.stages.push_back(s);
      plan.algorithmic_bytes += 0.5 * sweep_bytes;
      Gate a; a.kind = G_REFLECT; a.src_op = plan.gates[pending[0]].src_op;
      pending.erase(pending.begin());
      build_tile_stage(&a);
      continue;
    }
    // ---- multi-GPU: a non-diagonal target on a global physical bit needs a remap first
    {
      Gate g0 = to_phys(plan.gates[pending[0]]);
      uint64_t gt = g0.target_mask() & ~local_mask;
      if (gt) {
        int gbit = 63 - __builtin_clzll(gt);
        std::vector<int> logical_of(n);
        for (int b = 0; b < n; ++b) logical_of[perm[b]] = b;
        // swap with the local bit (among the top 8: large contiguous chunks) whose logical occupant is
        // needed latest as a non-diagonal target
        int best = -1; size_t best_next = 0;
        for (int cand = nl - 1; cand >= std::max(L, nl - 8) && cand >= 0; --cand) {
          if ((g0.target_mask() >> cand) & 1) continue;
          int lb = logical_of[cand];
          size_t next = pending.size() + 1;
          for (size_t i = 0; i < pending.size() && i < 2048; ++i)
            if ((plan.gates[pending[i]].target_mask() >> lb) & 1) { next = i; break; }
          if (best < 0 || next > best_next) { best = cand; best_next = next; }
        }
        if (best < 0) { plan.error = "no local qubit available for remap"; return QCB_ERR_INVALID; }
        Stage s; s.kind = S_EXCHANGE; s.gbit = gbit; s.lbit = best;
        plan.stages.push_back(s);
        plan.n_exchanges++;
        std::swap(perm[logical_of[gbit]], perm[logical_of[best]]);
        continue;
      }
    }
    build_tile_stage(nullptr);
  }

  plan.perm_out = perm;
  // ---- encode
  plan.words.clear();
  plan.stage_offsets.clear();
  plan.words.push_back(0x51434232ULL);                 // magic "QCB2"
  plan.words.push_back((uint64_t)plan.stages.size());
  plan.words.push_back((uint64_t)n);
  plan.words.push_back((uint64_t)nl);
  plan.n_rounds = 0;
  for (Stage& s : plan.stages) {
    plan.stage_offsets.push_back(plan.words.size());
    plan.words.push_back((uint64_t)s.kind);
    if (s.kind == S_TILE) { plan.words.push_back(0); encode_stage(cfg, s, plan.words); plan.n_rounds += s.rounds.size(); }
    else if (s.kind == S_EXCHANGE) { plan.words.push_back((uint64_t)s.gbit | ((uint64_t)s.lbit << 8)); }
    else { plan.words.push_back(0); }
  }
  return QCB_OK;
}

}  // namespace qcb
