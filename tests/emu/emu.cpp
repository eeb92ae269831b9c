// emu.cpp — HOST EMULATOR of the tile executor.  TEST INFRASTRUCTURE ONLY.
//
// It runs the scheduler's encoded program (plan.cpp) with the very same interpreter code the CUDA
// kernel uses (csrc/tile_core.h), replacing the CTA/thread grid by plain loops: every __syncthreads()
// boundary of k_tile_stage becomes a loop boundary here.  This lets the CPU test-suite check the
// scheduler, the ext-index encoding, the swizzle/lane mapping and the op interpreters against the
// oracle without a GPU.  It is built into tests/emu/libqcbemu.so by the tests themselves; nothing in
// qclojure_b200/ or libqcb200.so links, loads or falls back to it.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../qclojure_b200/csrc/plan.h"
#include "../../qclojure_b200/csrc/tile_core.h"

using namespace qcb;

struct Emu {
  Plan plan;
  std::string err;
};

extern "C" {

static uint64_t g_next_support = ~0ULL;
// support of the input state for the plans created from now on (plan.h: Plan::support_in; ~0 = unknown, 0 = |0...0>)
void emu_set_support(uint64_t s) { g_next_support = s; }

int emu_create(const qcb_config* cfg, const qcb_op* ops, uint64_t n_ops, const int32_t* perm_in, Emu** out) {
  Emu* e = new Emu();
  e->plan.cfg = config_from(*cfg);
  e->plan.support_in = g_next_support;
  int rc = lower_ops(e->plan.cfg, ops, n_ops, e->plan.gates, e->err);
  if (rc == QCB_OK) {
    std::vector<int> perm;
    if (perm_in) perm.assign(perm_in, perm_in + cfg->n_qubits);
    rc = schedule(e->plan, perm);
    if (rc != QCB_OK) e->err = e->plan.error;
  }
  *out = e;
  return rc;
}

const char* emu_error(Emu* e) { return e->err.c_str(); }
void emu_destroy(Emu* e) { delete e; }
uint64_t emu_num_stages(Emu* e) { return e->plan.stages.size(); }
int emu_stage_kind(Emu* e, uint64_t i) { return e->plan.stages[i].kind; }
void emu_stage_exchange(Emu* e, uint64_t i, int* gbit, int* lbit) { *gbit = e->plan.stages[i].gbit; *lbit = e->plan.stages[i].lbit; }
// S_GROVER: marked local indices of this rank (out, capacity 8), returns their number; *needs_sum = a diffusion follows
int emu_stage_grover(Emu* e, uint64_t i, uint64_t* marked, int* needs_sum) {
  const Stage& st = e->plan.stages[i];
  *needs_sum = st.needs_sum ? 1 : 0;
  for (size_t k = 0; k < st.marked.size() && k < 8; ++k) marked[k] = st.marked[k];
  return (int)st.marked.size();
}
uint64_t emu_num_rounds(Emu* e) { return e->plan.n_rounds; }
uint64_t emu_num_gates(Emu* e) { return e->plan.gates.size(); }
double emu_algorithmic_bytes(Emu* e) { return e->plan.algorithmic_bytes; }
double emu_unfused_bytes(Emu* e) { return e->plan.unfused_bytes; }
void emu_perm_out(Emu* e, int32_t* out) { for (size_t b = 0; b < e->plan.perm_out.size(); ++b) out[b] = e->plan.perm_out[b]; }
uint64_t emu_stage_num_gates(Emu* e, uint64_t i) { return e->plan.stages[i].src_gates.size(); }
uint64_t emu_stage_num_rounds(Emu* e, uint64_t i) { return e->plan.stages[i].rounds.size(); }
double emu_stage_fraction(Emu* e, uint64_t i) { return e->plan.stages[i].sweep_fraction; }
// per-round detail for plan analysis: number of gates, slot bits (as a mask over tile-local positions), condition bits
uint64_t emu_round_info(Emu* e, uint64_t si, uint64_t r, uint64_t* n_gates, uint64_t* slot_mask, uint64_t* n_cond) {
  const Round& rd = e->plan.stages[si].rounds[r];
  *n_gates = rd.gates.size();
  uint64_t m = 0;
  for (int p : rd.slot_pos) m |= 1ULL << p;
  *slot_mask = m;
  *n_cond = rd.cond_pos.size();
  return rd.dmma ? 1 : 0;
}
uint64_t emu_stage_tile_mask(Emu* e, uint64_t si) {
  uint64_t m = 0;
  for (int p : e->plan.stages[si].tile_pos) m |= 1ULL << p;
  return m;
}

// bank-conflict audit of one stage: worst number of distinct 16-byte bank-group collisions in any
// quarter-warp of any round (1 = conflict free)
int emu_stage_max_conflict(Emu* e, uint64_t si, int nthreads) {
  const uint64_t* st = e->plan.words.data() + e->plan.stage_offsets[si] + 2;
  StageCtx sc; decode_stage(st, sc);
  int worst = 1;
  for (uint32_t r = 0; r < sc.n_rounds; ++r) {
    if (round_kind(st, r) == 2u || round_kind(st, r) == 3u) {
      // 16-byte accesses are served per quarter-warp (8 lanes x 16 B = 128 B): count lanes per 16-byte bank group
      K3Ctx c; decode_k3(st, r, c);
      const bool pair = round_kind(st, r) == 3u;
      // the stores of a direct-store round go to global memory, not to the shared tile
      const int n_which = ((st[41] & T_FLAG_DIRECT_STORE) && r + 1u == sc.n_rounds) ? 2 : 4;
      for (int quarter = 0; quarter < 4; ++quarter)
        for (int which = 0; which < n_which; ++which) {
          int cnt[8] = {0};
          for (uint32_t l = 0; l < 8; ++l) {
            uint32_t e[4];
            if (pair) k3x_lane_entry(c, quarter * 8 + l, e); else k3_lane_entry(c, quarter * 8 + l, e);
            cnt[(e[which] >> 4) & 7]++;
          }
          for (int b = 0; b < 8; ++b) if (cnt[b] > worst) worst = cnt[b];
        }
      continue;
    }
    if (round_kind(st, r) == 1u) {
      // 8-byte accesses are served per half-warp (16 lanes x 8 B = 128 B): count lanes per 8-byte bank pair
      DmmaCtx c; decode_dmma(st, r, c);
      for (int half = 0; half < 2; ++half) {
        for (int which = 0; which < 8; ++which) {      // 4 load registers + 4 store registers
          int cnt[16] = {0};
          for (uint32_t l = 0; l < 16; ++l) {
            uint32_t lane = half * 16 + l, Pl[4], Ps[4], cl, cs;
            dmma_lane_setup(c, lane, Pl, Ps, cl, cs);
            uint32_t dw = which < 4 ? 2u * Pl[which] + cl : 2u * Ps[which - 4] + cs;   // index in doubles
            cnt[dw & 15]++;
          }
          for (int b = 0; b < 16; ++b) if (cnt[b] > worst) worst = cnt[b];
        }
      }
      continue;
    }
    RoundCtx rc; decode_round(st, r, rc);
    uint32_t ngroups = 1u << (sc.m - rc.r);
    for (uint32_t g0 = 0; g0 < ngroups && g0 < (uint32_t)nthreads; g0 += 8) {
      for (uint32_t s = 0; s < (1u << rc.r); ++s) {
        uint32_t off = 0;
        for (uint32_t j = 0; j < rc.r; ++j) if ((s >> j) & 1) off |= 1u << rc.slot_pos[j];
        int cnt[8] = {0};
        for (uint32_t l = 0; l < 8 && g0 + l < ngroups; ++l) cnt[swz(group_idx0(rc, g0 + l) | off, sc.c) & 7]++;
        for (int b = 0; b < 8; ++b) if (cnt[b] > worst) worst = cnt[b];
      }
    }
  }
  return worst;
}

// Software model of mma.sync.m16n8k16.f64 (PTX ISA fragment layouts, the same ones kernels.cu relies on) and of
// kernels.cu:dmma_round_device, warp by warp.
static void emu_dmma_round(double2* tile, const uint64_t* st, uint32_t r, uint64_t ext_hi, uint32_t m, int nthreads) {
  DmmaCtx c; decode_dmma(st, r, c);
  const uint32_t NW = nthreads / 32;
  const uint32_t nbatch = 1u << (c.n_grp - 3u);
  const uint32_t per = nbatch >= NW ? nbatch / NW : 1u;
  unsigned char* tb = reinterpret_cast<unsigned char*>(tile);
  const double* mats = reinterpret_cast<const double*>(st + c.mat_off);
  // the kernel's per-launch tables (tile_core.h: dmma_lane_entry / dmma_batch_entry), byte-offset addressing
  uint32_t lt[32][8];
  for (uint32_t lane = 0; lane < 32; ++lane) dmma_lane_entry(c, lane, lt[lane]);
  const uint32_t var_hi = dmma_variant_hi(c, ext_hi, m);
  for (uint32_t warp = 0; warp < NW; ++warp) {
    for (uint32_t b = 0; b < per; ++b) {
      const uint32_t bidx = warp * per + b;
      if (bidx >= nbatch) break;
      const uint32_t e = dmma_batch_entry(c, bidx, m);
      const uint32_t X = e & DMMA_BATCH_OFF_MASK, var = var_hi | (e >> 20);
      if (var != dmma_variant(c, dmma_batch_base(c, bidx), ext_hi, m)) std::abort();
      double A[32][8], B[32][4], D[32][4];
      for (uint32_t lane = 0; lane < 32; ++lane) {
        for (int i = 0; i < 8; ++i) A[lane][i] = mats[((size_t)var * 8 + i) * 32 + lane];
        for (int v = 0; v < 4; ++v) std::memcpy(&B[lane][v], tb + (X ^ lt[lane][v]), 8);
      }
      double Am[16][16], Bm[16][8], Dm[16][8];
      for (uint32_t lane = 0; lane < 32; ++lane) {
        for (int i = 0; i < 8; ++i) Am[lane / 4 + 8 * (i & 1)][lane % 4 + 4 * (i >> 1)] = A[lane][i];
        for (int v = 0; v < 4; ++v) Bm[lane % 4 + 4 * v][lane / 4] = B[lane][v];
      }
      for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) { double acc = 0; for (int k = 0; k < 16; ++k) acc += Am[i][k] * Bm[k][j]; Dm[i][j] = acc; }
      for (uint32_t lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 4; ++i) D[lane][i] = Dm[lane / 4 + 8 * (i >> 1)][2 * (lane % 4) + (i & 1)];
      for (uint32_t lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 4; ++i) std::memcpy(tb + (X ^ lt[lane][4 + i]), &D[lane][i], 8);
    }
  }
}

// Software model of the three-product round (tile_core.h: K3Ctx; kernels.cu: k3_round_run), warp by warp, with the
// mma.m8n8k4 .f64 fragment layouts: K = P Br, Re = K + N (Br + Bi), Im = K + R (Bi - Br).
// direct != nullptr: the round is the last one of a sweep with T_FLAG_DIRECT_STORE - results go to global memory at
// direct[goff(tile-local index)] exactly like the kernel computes the address (XOR of the batch's and the lane's offsets)
static void emu_k3_round(double2* tile, const uint64_t* st, uint32_t r, uint64_t ext_hi, uint32_t m, int nthreads,
                         double2* direct = nullptr, const StageCtx* scp = nullptr) {
  K3Ctx c; decode_k3(st, r, c);
  auto goff = [&](uint32_t idx) -> uint64_t { return hi_offset(st, *scp, idx >> scp->L) + (idx & ((1u << scp->L) - 1u)); };
  const uint32_t NW = nthreads / 32;
  const uint32_t nbatch = 1u << (c.n_grp - 3u);
  const uint32_t per = nbatch >= NW ? nbatch / NW : 1u;
  unsigned char* tb = reinterpret_cast<unsigned char*>(tile);
  const double* mats = reinterpret_cast<const double*>(st + c.mat_off);
  uint32_t lt[32][4];
  for (uint32_t lane = 0; lane < 32; ++lane) k3_lane_entry(c, lane, lt[lane]);
  const uint32_t var_hi = k3_variant_hi(c, ext_hi, m);
  for (uint32_t warp = 0; warp < NW; ++warp) {
    for (uint32_t b = 0; b < per; ++b) {
      const uint32_t bidx = warp * per + b;
      if (bidx >= nbatch) break;
      const uint32_t e = k3_batch_entry(c, bidx, m);
      const uint32_t X = e & DMMA_BATCH_OFF_MASK, var = var_hi | (e >> 20);
      double Pm[8][8], Nm[8][8], Rm[8][8], Br[8][8], Bi[8][8];
      for (uint32_t lane = 0; lane < 32; ++lane) {
        for (int s = 0; s < 2; ++s) {
          const int row = lane / 4, col = lane % 4 + 4 * s;
          Pm[row][col] = mats[(size_t)var * K3_FRAG_DOUBLES + (0 + s) * 32 + lane];
          Nm[row][col] = mats[(size_t)var * K3_FRAG_DOUBLES + (2 + s) * 32 + lane];
          Rm[row][col] = mats[(size_t)var * K3_FRAG_DOUBLES + (4 + s) * 32 + lane];
          double2 a; std::memcpy(&a, tb + (X ^ lt[lane][s]), 16);
          Br[lane % 4 + 4 * s][lane / 4] = a.x; Bi[lane % 4 + 4 * s][lane / 4] = a.y;
        }
      }
      if (c.far_n[0]) {      // far phases: every row of the block scaled by the tile's phase for that output pattern
        double fs[4];
        far_sums(st + c.far_off[0], c.far_n[0], ext_hi, fs);
        for (int row = 0; row < 8; ++row) {
          const double ang = far_angle(fs, k3_pattern_index((uint32_t)row, c.mmap));
          const double dr = std::cos(ang), di = std::sin(ang);
          for (int col = 0; col < 8; ++col) far_scale(Pm[row][col], Nm[row][col], Rm[row][col], dr, di);
        }
      }
      if (c.far_n[2]) {      // ... and, for phases that precede the block, every column (input pattern)
        double fs[4];
        far_sums(st + c.far_off[2], c.far_n[2], ext_hi, fs);
        for (int col = 0; col < 8; ++col) {
          const double ang = far_angle(fs, k3_pattern_index((uint32_t)col, c.kmap));
          const double dr = std::cos(ang), di = std::sin(ang);
          for (int row = 0; row < 8; ++row) far_scale(Pm[row][col], Nm[row][col], Rm[row][col], dr, di);
        }
      }
      double Re[8][8], Im[8][8];
      for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
        double k = 0, re = 0, im = 0;
        for (int t = 0; t < 8; ++t) k += Pm[i][t] * Br[t][j];
        re = k; im = k;
        for (int t = 0; t < 8; ++t) { re += Nm[i][t] * (Br[t][j] + Bi[t][j]); im += Rm[i][t] * (Bi[t][j] - Br[t][j]); }
        Re[i][j] = re; Im[i][j] = im;
      }
      for (uint32_t lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 2; ++i) {
          const double2 o{Re[lane / 4][2 * (lane % 4) + i], Im[lane / 4][2 * (lane % 4) + i]};
          if (direct) direct[goff(k3_batch_base(c, bidx)) ^ goff(k3_lane_store_index(c, lane, (uint32_t)i))] = o;
          else std::memcpy(tb + (X ^ lt[lane][2 + i]), &o, 16);
        }
    }
  }
}

// Software model of a paired round (tile_core.h "paired rounds"; kernels.cu: k3x_round_run), lane by lane: the first block's
// D registers are handed to the second block as its B registers exactly as the kernel does it, so the free transposition the
// m8n8k4 fragment layouts provide is what is being tested here.
static void emu_k3x_round(double2* tile, const uint64_t* st, uint32_t r, uint64_t ext_hi, uint32_t m, int nthreads) {
  K3Ctx c; decode_k3(st, r, c);
  const uint32_t NW = nthreads / 32;
  const uint32_t nbatch = 1u << (c.n_grp - 3u);
  const uint32_t per = nbatch >= NW ? nbatch / NW : 1u;
  unsigned char* tb = reinterpret_cast<unsigned char*>(tile);
  const double* mats = reinterpret_cast<const double*>(st + c.mat_off);
  uint32_t lt[32][4];
  for (uint32_t lane = 0; lane < 32; ++lane) k3x_lane_entry(c, lane, lt[lane]);
  const uint32_t var_hi = k3_variant_hi(c, ext_hi, m);
  // D = A (8x8, fragment registers a0 / a1 per lane = k-steps 0 / 1) * B (8x8, registers b0 / b1 per lane) + C, all in the
  // per-lane register layouts of mma.m8n8k4 .f64
  auto mma = [](const double (&a0)[32], const double (&a1)[32], const double (&b0)[32], const double (&b1)[32], const double (&c0)[32],
                const double (&c1)[32], double (&d0)[32], double (&d1)[32]) {
    double Am[8][8], Bm[8][8];
    for (int lane = 0; lane < 32; ++lane) {
      Am[lane / 4][lane % 4] = a0[lane]; Am[lane / 4][lane % 4 + 4] = a1[lane];
      Bm[lane % 4][lane / 4] = b0[lane]; Bm[lane % 4 + 4][lane / 4] = b1[lane];
    }
    double o0[32], o1[32];
    for (int lane = 0; lane < 32; ++lane) {
      const int row = lane / 4, col = 2 * (lane % 4);
      double x = c0[lane], y = c1[lane];
      for (int t = 0; t < 8; ++t) { x += Am[row][t] * Bm[t][col]; y += Am[row][t] * Bm[t][col + 1]; }
      o0[lane] = x; o1[lane] = y;
    }
    for (int lane = 0; lane < 32; ++lane) { d0[lane] = o0[lane]; d1[lane] = o1[lane]; }
  };
  for (uint32_t warp = 0; warp < NW; ++warp) {
    for (uint32_t b = 0; b < per; ++b) {
      const uint32_t bidx = warp * per + b;
      if (bidx >= nbatch) break;
      const uint32_t e = k3_batch_entry(c, bidx, m);
      const uint32_t X = e & DMMA_BATCH_OFF_MASK, var = var_hi | (e >> 20);
      double A[12][32], r0[32], i0[32], r1[32], i1[32], zero[32] = {};
      double fs1[4] = {0, 0, 0, 0}, fs2[4] = {0, 0, 0, 0};
      if (c.far_n[0]) far_sums(st + c.far_off[0], c.far_n[0], ext_hi, fs1);
      if (c.far_n[1]) far_sums(st + c.far_off[1], c.far_n[1], ext_hi, fs2);
      double fp1[4] = {0, 0, 0, 0}, fp2[4] = {0, 0, 0, 0};
      if (c.far_n[2]) far_sums(st + c.far_off[2], c.far_n[2], ext_hi, fp1);
      if (c.far_n[3]) far_sums(st + c.far_off[3], c.far_n[3], ext_hi, fp2);
      for (uint32_t lane = 0; lane < 32; ++lane) {
        for (int i = 0; i < 12; ++i) A[i][lane] = mats[(size_t)var * K3X_FRAG_DOUBLES + i * 32 + lane];
        // far phases: a lane's fragments all belong to the row lane / 4 of their block
        if (c.far_n[0]) {
          const double ang = far_angle(fs1, k3_pattern_index(lane / 4, c.mmap));
          for (int s2 = 0; s2 < 2; ++s2) far_scale(A[0 + s2][lane], A[2 + s2][lane], A[4 + s2][lane], std::cos(ang), std::sin(ang));
        }
        if (c.far_n[1]) {
          const double ang = far_angle(fs2, k3_pattern_index(lane / 4, c.mmap2));
          for (int s2 = 0; s2 < 2; ++s2) far_scale(A[6 + s2][lane], A[8 + s2][lane], A[10 + s2][lane], std::cos(ang), std::sin(ang));
        }
        // phases that precede a block: column scaling, a lane's fragment register s belongs to column lane % 4 + 4 s
        for (int s2 = 0; s2 < 2; ++s2) {
          const uint32_t kcol = lane % 4 + 4 * s2;
          if (c.far_n[2]) {
            const double ang = far_angle(fp1, k3_pattern_index(kcol, c.kmap));
            far_scale(A[0 + s2][lane], A[2 + s2][lane], A[4 + s2][lane], std::cos(ang), std::sin(ang));
          }
          if (c.far_n[3]) {
            const double ang = far_angle(fp2, k3x_hw_k_to_group(kcol));
            far_scale(A[6 + s2][lane], A[8 + s2][lane], A[10 + s2][lane], std::cos(ang), std::sin(ang));
          }
        }
        double2 a; std::memcpy(&a, tb + (X ^ lt[lane][0]), 16); r0[lane] = a.x; i0[lane] = a.y;
        std::memcpy(&a, tb + (X ^ lt[lane][1]), 16); r1[lane] = a.x; i1[lane] = a.y;
      }
      for (int blk = 0; blk < 2; ++blk) {
        double s0[32], s1[32], d0[32], d1[32], K0[32], K1[32], x0[32], x1[32], y0[32], y1[32];
        for (int lane = 0; lane < 32; ++lane) { s0[lane] = r0[lane] + i0[lane]; d0[lane] = i0[lane] - r0[lane]; s1[lane] = r1[lane] + i1[lane]; d1[lane] = i1[lane] - r1[lane]; }
        const int o = 6 * blk;
        mma(A[o + 0], A[o + 1], r0, r1, zero, zero, K0, K1);
        mma(A[o + 2], A[o + 3], s0, s1, K0, K1, x0, x1);
        mma(A[o + 4], A[o + 5], d0, d1, K0, K1, y0, y1);
        // results (x = Re, y = Im) of columns 2(lane%4) + {0, 1} become the operands of k-steps 0, 1
        for (int lane = 0; lane < 32; ++lane) { r0[lane] = x0[lane]; r1[lane] = x1[lane]; i0[lane] = y0[lane]; i1[lane] = y1[lane]; }
      }
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const double2 o0{r0[lane], i0[lane]}, o1{r1[lane], i1[lane]};
        std::memcpy(tb + (X ^ lt[lane][2]), &o0, 16);
        std::memcpy(tb + (X ^ lt[lane][3]), &o1, 16);
      }
    }
  }
}

// Run one S_TILE stage on this rank's local slice.
int emu_run_tile_stage(Emu* e, uint64_t si, double* state, const double* dev_vals, int nthreads) {
  const Stage& S = e->plan.stages[si];
  if (S.kind != S_TILE) return QCB_ERR_INVALID;
  const uint64_t* st = e->plan.words.data() + e->plan.stage_offsets[si] + 2;
  StageCtx sc; decode_stage(st, sc);
  double2* gs = reinterpret_cast<double2*>(state);
  const uint32_t tile_n = 1u << sc.m;
  std::vector<double2> tile(tile_n);
  const uint32_t nb = sc.n_local - sc.m;
  const uint64_t tmask = (nb >= 64) ? ~0ULL : ((1ULL << nb) - 1ULL);
  const uint64_t n_active = (1ULL << nb) >> __builtin_popcountll(sc.skip_mask & tmask);
  // rank-level skip (the part of the condition that lives in the rank bits)
  if (((sc.ext_hi_base & sc.skip_mask & ~tmask) != (sc.skip_val & ~tmask))) return QCB_OK;
  for (uint64_t a = 0; a < n_active; ++a) {
    const uint64_t t = active_to_tile(sc, a);
    const uint64_t ext_hi = sc.ext_hi_base | t;
    const uint64_t base = tile_base(st, sc, t);
    for (uint32_t i = 0; i < tile_n; ++i)
      tile[swz(i, sc.c)] = gs[base + hi_offset(st, sc, i >> sc.L) + (i & ((1u << sc.L) - 1u))];
    for (uint32_t r = 0; r < sc.n_rounds; ++r) {
      if (round_kind(st, r) == 1u) { emu_dmma_round(tile.data(), st, r, ext_hi, sc.m, nthreads); continue; }
      if (round_kind(st, r) == 2u) {
        const bool direct = (st[41] & T_FLAG_DIRECT_STORE) && r + 1u == sc.n_rounds;
        emu_k3_round(tile.data(), st, r, ext_hi, sc.m, nthreads, direct ? gs + base : nullptr, &sc);
        continue;
      }
      if (round_kind(st, r) == 3u) { emu_k3x_round(tile.data(), st, r, ext_hi, sc.m, nthreads); continue; }
      RoundCtx rc; decode_round(st, r, rc);
      for (uint32_t tid = 0; tid < (uint32_t)nthreads; ++tid) {
        switch (rc.r) {
          case 0: run_round_thread<0>(tile.data(), rc, sc.m, sc.c, ext_hi, tid, nthreads, dev_vals); break;
          case 1: run_round_thread<1>(tile.data(), rc, sc.m, sc.c, ext_hi, tid, nthreads, dev_vals); break;
          case 2: run_round_thread<2>(tile.data(), rc, sc.m, sc.c, ext_hi, tid, nthreads, dev_vals); break;
          default: run_round_thread<3>(tile.data(), rc, sc.m, sc.c, ext_hi, tid, nthreads, dev_vals); break;
        }
      }
    }
    if (st[41] & T_FLAG_DIRECT_STORE) continue;      // the last round has written the tile itself
    for (uint32_t i = 0; i < tile_n; ++i)
      gs[base + hi_offset(st, sc, i >> sc.L) + (i & ((1u << sc.L) - 1u))] = tile[swz(i, sc.c)];
  }
  return QCB_OK;
}

// S_SUM partial: sum of this rank's amplitudes
void emu_local_sum(const double* state, uint64_t count, double out[2]) {
  double re = 0, im = 0;
  for (uint64_t i = 0; i < count; ++i) { re += state[2 * i]; im += state[2 * i + 1]; }
  out[0] = re; out[1] = im;
}

uint32_t emu_swz(uint32_t i, uint32_t c) { return swz(i, c); }

uint64_t emu_program_words(Emu* e, uint64_t* out, uint64_t cap) {
  uint64_t n = e->plan.words.size();
  if (out) std::memcpy(out, e->plan.words.data(), sizeof(uint64_t) * (n < cap ? n : cap));
  return n;
}

}  // extern "C"
