// tile_core.h — the arithmetic core of the fused gate executor, shared verbatim between the CUDA
// kernel (kernels.cu: k_tile_stage) and the host emulator used by the CPU test-suite (tests/emu/).
//
// Data model.  One *stage* sweeps the local state once.  A CTA owns one *tile* of 2^m amplitudes:
// the m tile bits are the L lowest index bits (a contiguous, coalesced run of 2^L * 16 bytes) plus
// m-L arbitrary higher bits chosen by the scheduler.  The tile lives in shared memory as interleaved
// double2, XOR-swizzled so that 16-byte accesses of a quarter-warp fall into 8 distinct bank groups.
// A stage is a list of *rounds*; in a round every thread keeps 2^r amplitudes (r <= 3 "slot" bits) in
// registers and interprets the round's op list on them; rounds are separated by __syncthreads().
//
// The "ext" index of an amplitude = (rank | tile id | tile-local index): bit positions [0,m) are the
// tile-local bits, [m, n_local) the tile-id bits, [n_local, n_total) the rank bits.  Control masks and
// diagonal predicates are expressed in ext space by the scheduler (plan.cpp: to_ext).
#pragma once
#include <stdint.h>
#include <vector_types.h>   // double2 (plain header, usable from g++ as well)

#if defined(__CUDACC__)
#define QCB_HD __host__ __device__ __forceinline__
#else
#define QCB_HD inline
#endif

namespace qcb {

// must match plan.h (DevOpKind)
enum : uint32_t { TD_MAT1 = 0, TD_MAT2 = 1, TD_SWAPP = 2, TD_DMASK = 3, TD_DNEG = 4, TD_DPOP1 = 5, TD_AFFINE = 6,
                  TD_MAT1R = 7, TD_MAT1RI = 8, TD_PERMX = 9, TD_DENSE = 10 };
constexpr int T_OP_WORDS = 16, T_STAGE_WORDS = 48, T_ROUND_WORDS = 40;
// stage word [41] flags (plan.h): bit 1 = the last round (a three-product tensor-core round) stores its results straight to
// global memory from registers; the mover then never writes the tile back
constexpr uint64_t T_FLAG_DIRECT_STORE = 2;

// Shared-memory layout of a tile (16-byte units), parameter c = stage word [43].
// c = 0 (LSU mover, the default): every row bit >= 3 is XOR-folded onto the three chunk bits - any three index bits
// with distinct residues mod 3 give conflict-free access.
// c >= 3 (TMA mover): the tile is written by TMA tensor copies with the hardware 128-byte swizzle, one copy per
// contiguous run of 2^c amplitudes (c = number of contiguous low tile bits), i.e. 2^(c-3) rows of 128 bytes whose
// 16-byte chunk index is XORed with bits 7..9 of the shared-memory address.  Row index r = i >> 3; the run is placed at row address r ^ fold(r), where fold() XORs the row bits >= 3
// cyclically onto the row-address bits [c-3, 3) that a run does not fix itself.  Net effect: index bit p lands on
// chunk bit  p (p < 3),  p-3 (3 <= p < 6),  and for p >= 6:  (c-3) + (p-6) mod (6-c)  when c < 6, none otherwise.
// Accesses whose varying index bits fall on three distinct chunk bits are bank-conflict free (plan.cpp: chunk_class).
// The map is linear over GF(2): swz(a ^ b, c) == swz(a, c) ^ swz(b, c).
QCB_HD uint32_t swz_fold(uint32_t r, uint32_t c) {
  if (c >= 6) return 0;
  const uint32_t t = r >> 3;
  if (c <= 3) return (t ^ (t >> 3) ^ (t >> 6) ^ (t >> 9)) & 7u;
  if (c == 4) return ((t ^ (t >> 2) ^ (t >> 4) ^ (t >> 6) ^ (t >> 8)) & 3u) << 1;
  uint32_t x = t; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;      // c == 5: parity
  return (x & 1u) << 2;
}
QCB_HD uint32_t swz(uint32_t i, uint32_t c) {
  const uint32_t r = i >> 3, ra = r ^ swz_fold(r, c);
  return (ra << 3) | ((i & 7u) ^ (ra & 7u));
}

QCB_HD uint32_t insert_zero(uint32_t v, uint32_t pos) { return ((v >> pos) << (pos + 1)) | (v & ((1u << pos) - 1u)); }

QCB_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

QCB_HD double as_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; __builtin_memcpy(&d, &u, 8); return d;
#endif
}

#if defined(__CUDA_ARCH__)
#define QCB_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define QCB_FMA(a, b, c) ((a) * (b) + (c))
#endif

QCB_HD double2 cmul(double2 a, double2 b) {
  return double2{QCB_FMA(a.x, b.x, -(a.y * b.y)), QCB_FMA(a.x, b.y, a.y * b.x)};
}
// a*b + c*d as two chains of 1 mul + 3 fma
QCB_HD double2 cmul2(double2 a, double2 b, double2 c, double2 d) {
  double re = c.x * d.x;
  re = QCB_FMA(-c.y, d.y, re);
  re = QCB_FMA(a.x, b.x, re);
  re = QCB_FMA(-a.y, b.y, re);
  double im = c.x * d.y;
  im = QCB_FMA(c.y, d.x, im);
  im = QCB_FMA(a.x, b.y, im);
  im = QCB_FMA(a.y, b.x, im);
  return double2{re, im};
}

struct RoundCtx {
  uint32_t r, n_slots, n_lane, n_ins;
  uint32_t slot_pos[3], lane_pos[3], ins_pos[6];
  const uint64_t* ops;
};

QCB_HD void decode_round(const uint64_t* stage, uint32_t round_idx, RoundCtx& rc) {
  const uint64_t* w = stage + T_STAGE_WORDS + (uint64_t)round_idx * T_ROUND_WORDS;
  rc.r = (uint32_t)w[0]; rc.n_slots = (uint32_t)w[1]; rc.ops = stage + w[2]; rc.n_lane = (uint32_t)w[3];
  for (int j = 0; j < 3; ++j) { rc.slot_pos[j] = (uint32_t)w[4 + j]; rc.lane_pos[j] = (uint32_t)w[7 + j]; }
  rc.n_ins = (uint32_t)w[10];
  for (int j = 0; j < 6; ++j) rc.ins_pos[j] = (uint32_t)w[11 + j];
}

// group index g in [0, 2^(m-r)) -> tile-local index with all slot bits = 0
QCB_HD uint32_t group_idx0(const RoundCtx& rc, uint32_t g) {
  uint32_t lane = g & ((1u << rc.n_lane) - 1u);
  uint32_t v = g >> rc.n_lane;
  // fully unrolled with static indices so that RoundCtx stays in registers on the device
#pragma unroll
  for (uint32_t k = 0; k < 6; ++k) if (k < rc.n_ins) v = insert_zero(v, rc.ins_pos[k]);
#pragma unroll
  for (uint32_t j = 0; j < 3; ++j) if (j < rc.n_lane) v |= ((lane >> j) & 1u) << rc.lane_pos[j];
  return v;
}

// ---- op interpreters on 2^R register-resident amplitudes.
// The scheduler pre-splits every condition (controls / diagonal predicates) of an op into
//   sel      8-bit bitmap over the slot patterns s for which the slot-bit part of the condition holds
//            (for pair/quad ops: indexed by the base pattern, target bits = 0),
//   loc      mask|val on the tile-local NON-slot bits (32-bit, compared against idx0 once per group),
//   hi       mask/val on the tile-id + rank bits (64-bit, uniform over the CTA's current tile),
// so the per-amplitude work is one bit test instead of 64-bit mask arithmetic.
template <int R, int MODE>   // MODE 0: general complex 2x2; 1: real 2x2; 2: real diagonal, imaginary off-diagonal
QCB_HD void op_mat1(double2 (&a)[1 << R], uint32_t j, uint32_t sel, const uint64_t* mw) {
  double c[8];
#pragma unroll
  for (int k = 0; k < (MODE == 0 ? 8 : 4); ++k) c[k] = as_double(mw[k]);
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    if ((uint32_t)jj != j) continue;
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) {
      if (s & (1 << jj)) continue;
      const int s1 = s | (1 << jj);
      if (sel & (1u << s)) {
        const double2 a0 = a[s], a1 = a[s1];
        if (MODE == 0) {
          a[s] = cmul2(double2{c[0], c[1]}, a0, double2{c[2], c[3]}, a1);
          a[s1] = cmul2(double2{c[4], c[5]}, a0, double2{c[6], c[7]}, a1);
        } else if (MODE == 1) {      // m00 m01 m10 m11 real
          a[s] = double2{QCB_FMA(c[0], a0.x, c[1] * a1.x), QCB_FMA(c[0], a0.y, c[1] * a1.y)};
          a[s1] = double2{QCB_FMA(c[2], a0.x, c[3] * a1.x), QCB_FMA(c[2], a0.y, c[3] * a1.y)};
        } else {                     // m00 = c0, m01 = i c1, m10 = i c2, m11 = c3
          a[s] = double2{QCB_FMA(c[0], a0.x, -(c[1] * a1.y)), QCB_FMA(c[0], a0.y, c[1] * a1.x)};
          a[s1] = double2{QCB_FMA(c[3], a1.x, -(c[2] * a0.y)), QCB_FMA(c[3], a1.y, c[2] * a0.x)};
        }
      }
    }
  }
}

template <int R>
QCB_HD void op_permx(double2 (&a)[1 << R], uint32_t j, uint32_t sel) {
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    if ((uint32_t)jj != j) continue;
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) {
      if (s & (1 << jj)) continue;
      const int s1 = s | (1 << jj);
      if (sel & (1u << s)) { const double2 t = a[s]; a[s] = a[s1]; a[s1] = t; }
    }
  }
}

template <int R>
QCB_HD void op_swapp(double2 (&a)[1 << R], uint32_t j0, uint32_t j1, uint32_t sel, const uint64_t* mw) {
  const double2 ph{as_double(mw[0]), as_double(mw[1])};
  const bool unit = (ph.x == 1.0 && ph.y == 0.0);
#pragma unroll
  for (int p0 = 0; p0 < R; ++p0) {
#pragma unroll
    for (int p1 = p0 + 1; p1 < R; ++p1) {
      if ((uint32_t)p0 != j0 || (uint32_t)p1 != j1) continue;
#pragma unroll
      for (int s = 0; s < (1 << R); ++s) {
        if (((s >> p0) & 1) || ((s >> p1) & 1)) continue;     // base pattern: both target bits 0
        const int u = s | (1 << p0), v = s | (1 << p1);       // the two patterns that differ
        if (sel & (1u << s)) {
          double2 x = a[u], y = a[v];
          if (!unit) { x = cmul(ph, x); y = cmul(ph, y); }
          a[u] = y; a[v] = x;
        }
      }
    }
  }
}

template <int R>
QCB_HD void op_mat2(double2 (&a)[1 << R], uint32_t j0, uint32_t j1, uint32_t sel, const uint64_t* mw) {
#pragma unroll
  for (int p0 = 0; p0 < R; ++p0) {
#pragma unroll
    for (int p1 = p0 + 1; p1 < R; ++p1) {
      if ((uint32_t)p0 != j0 || (uint32_t)p1 != j1) continue;
#pragma unroll
      for (int s = 0; s < (1 << R); ++s) {
        if (((s >> p0) & 1) || ((s >> p1) & 1)) continue;
        if (!(sel & (1u << s))) continue;
        const int i0 = s, i1 = s | (1 << p0), i2 = s | (1 << p1), i3 = s | (1 << p0) | (1 << p1);
        const double2 v0 = a[i0], v1 = a[i1], v2 = a[i2], v3 = a[i3];
        double2 o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint64_t* row = mw + r * 8;
          double2 c0{as_double(row[0]), as_double(row[1])}, c1{as_double(row[2]), as_double(row[3])};
          double2 c2{as_double(row[4]), as_double(row[5])}, c3{as_double(row[6]), as_double(row[7])};
          double2 t0 = cmul2(c0, v0, c1, v1), t1 = cmul2(c2, v2, c3, v3);
          o[r] = double2{t0.x + t1.x, t0.y + t1.y};
        }
        a[i0] = o[0]; a[i1] = o[1]; a[i2] = o[2]; a[i3] = o[3];
      }
    }
  }
}

// fused dense 2^R x 2^R block on all slot bits: out[row] = sum_col M[row][col] * a[col]
template <int R>
QCB_HD void op_dense(double2 (&a)[1 << R], const uint64_t* mw) {
  constexpr int D = 1 << R;
  double2 o[D];
#pragma unroll
  for (int row = 0; row < D; ++row) {
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int col = 0; col < D; ++col) {
      const double mr = as_double(mw[2 * (row * D + col)]), mi = as_double(mw[2 * (row * D + col) + 1]);
      re = QCB_FMA(mr, a[col].x, re); re = QCB_FMA(-mi, a[col].y, re);
      im = QCB_FMA(mr, a[col].y, im); im = QCB_FMA(mi, a[col].x, im);
    }
    o[row] = double2{re, im};
  }
#pragma unroll
  for (int row = 0; row < D; ++row) a[row] = o[row];
}

template <int R>
QCB_HD void apply_ops(double2 (&a)[1 << R], const uint64_t* ops, uint32_t n_slots, uint32_t idx0, uint64_t ext_hi,
                      uint32_t m, const uint32_t (&off)[1 << R], const double* dev_vals) {
  for (uint32_t s = 0; s < n_slots;) {
    const uint64_t* w = ops + (uint64_t)s * T_OP_WORDS;
    const uint64_t hdr = w[0], loc = w[1];
    const uint32_t lo = (uint32_t)hdr, sel = (uint32_t)(hdr >> 32) & 0xffu;
    const uint32_t kind = lo & 0xffu, j0 = (lo >> 8) & 0xffu, j1 = (lo >> 16) & 0xffu, ns = (lo >> 24) & 0xffu;
    s += (ns ? ns : 1u);
    if ((idx0 & (uint32_t)loc) != (uint32_t)(loc >> 32)) continue;
    if ((ext_hi & w[2]) != w[3]) continue;
    switch (kind) {
      case TD_MAT1: op_mat1<R, 0>(a, j0, sel, w + 4); break;
      case TD_MAT1R: op_mat1<R, 1>(a, j0, sel, w + 4); break;
      case TD_MAT1RI: op_mat1<R, 2>(a, j0, sel, w + 4); break;
      case TD_PERMX: op_permx<R>(a, j0, sel); break;
      case TD_DENSE: op_dense<R>(a, w + 4); break;
      case TD_MAT2: op_mat2<R>(a, j0, j1, sel, w + 4); break;
      case TD_SWAPP: op_swapp<R>(a, j0, j1, sel, w + 4); break;
      case TD_DMASK: {
        const double2 ph{as_double(w[4]), as_double(w[5])};
#pragma unroll
        for (int k = 0; k < (1 << R); ++k)
          if (sel & (1u << k)) a[k] = cmul(a[k], ph);
        break;
      }
      case TD_DNEG: {
#pragma unroll
        for (int k = 0; k < (1 << R); ++k)
          if (sel & (1u << k)) a[k] = double2{-a[k].x, -a[k].y};
        break;
      }
      case TD_DPOP1: {   // phase where exactly one bit of the (ext-space) mask is set: rydberg-blockade
        const uint64_t mask = w[4];
        const double2 ph{as_double(w[6]), as_double(w[7])};
        const uint64_t e0 = (ext_hi << m) | idx0;
#pragma unroll
        for (int k = 0; k < (1 << R); ++k)
          if (popc64((e0 | off[k]) & mask) == 1) a[k] = cmul(a[k], ph);
        break;
      }
      case TD_AFFINE: {   // a' = alpha * a + beta, (alpha, beta) produced on the device by a reduction
        const double* v = dev_vals + 4 * w[4];
        const double2 al{v[0], v[1]}, be{v[2], v[3]};
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) { double2 t = cmul(al, a[k]); a[k] = double2{t.x + be.x, t.y + be.y}; }
        break;
      }
      default: break;
    }
  }
}

// One thread's share of a round: groups g = tid, tid+T, ... ; tile = swizzled shared-memory tile.
template <int R>
QCB_HD void run_round_thread(double2* tile, const RoundCtx& rc, uint32_t m, uint32_t c, uint64_t ext_hi, uint32_t tid,
                             uint32_t nthreads, const double* dev_vals) {
  uint32_t off[1 << R], soff[1 << R];
#pragma unroll
  for (int s = 0; s < (1 << R); ++s) {
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) if ((s >> j) & 1) o |= 1u << rc.slot_pos[j];
    off[s] = o;
    soff[s] = swz(o, c);
  }
  const uint32_t ngroups = 1u << (m - R);
  for (uint32_t g = tid; g < ngroups; g += nthreads) {
    const uint32_t idx0 = group_idx0(rc, g), s0 = swz(idx0, c);        // swz is linear: swz(idx0 | off) = s0 ^ soff
    double2 a[1 << R];
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) a[s] = tile[s0 ^ soff[s]];
    apply_ops<R>(a, rc.ops, rc.n_slots, idx0, ext_hi, m, off, dev_vals);
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) tile[s0 ^ soff[s]] = a[s];
  }
}

// ---- tensor-core ("dmma") rounds: the round is one of 2^k dense 16x16 real matrices applied to the 16 reals
// (8 slot patterns x re/im) of every group with mma.sync.m16n8k16.f64; 8 groups (the lane bits) form one MMA.
struct DmmaCtx {
  uint32_t n_grp, k, j_load, j_store, c;     // c: run bits of the stage (layout parameter of swz)
  uint32_t slot_pos[3], grp_pos[10], cond_pos[4];
  uint64_t mat_off;
};

QCB_HD uint32_t round_kind(const uint64_t* stage, uint32_t round_idx) {
  return (uint32_t)stage[T_STAGE_WORDS + (uint64_t)round_idx * T_ROUND_WORDS + 17];
}

QCB_HD void decode_dmma(const uint64_t* stage, uint32_t round_idx, DmmaCtx& c) {
  const uint64_t* w = stage + T_STAGE_WORDS + (uint64_t)round_idx * T_ROUND_WORDS;
  c.c = (uint32_t)stage[43];
  c.mat_off = w[2]; c.n_grp = (uint32_t)w[18]; c.k = (uint32_t)w[29]; c.j_load = (uint32_t)w[34]; c.j_store = (uint32_t)w[35];
  for (int j = 0; j < 3; ++j) c.slot_pos[j] = (uint32_t)w[4 + j];
  for (int j = 0; j < 10; ++j) c.grp_pos[j] = (uint32_t)w[19 + j];
  for (int j = 0; j < 4; ++j) c.cond_pos[j] = (uint32_t)w[30 + j];
}

// k-index / m-index of the MMA -> tile-local offset of the slot pattern: bit 1 = slot jsel, bits 2,3 = the two
// remaining slots in ascending order (bit 0 is the re/im component).  Must match plan.cpp:build_dmma_round.
QCB_HD uint32_t dmma_pattern_offset(const DmmaCtx& c, uint32_t idx, uint32_t jsel) {
  uint32_t rem0 = (jsel == 0) ? 1u : 0u, rem1 = (jsel == 2) ? 1u : 2u;
  uint32_t o = ((idx >> 1) & 1u) << (jsel == 0 ? c.slot_pos[0] : (jsel == 1 ? c.slot_pos[1] : c.slot_pos[2]));
  o |= ((idx >> 2) & 1u) << (rem0 == 0 ? c.slot_pos[0] : c.slot_pos[1]);
  o |= ((idx >> 3) & 1u) << (rem1 == 1 ? c.slot_pos[1] : c.slot_pos[2]);
  return o;
}

// tile-local offset of the lane-bit pattern n (column of the MMA = group n of the batch)
QCB_HD uint32_t dmma_lane_offset(const DmmaCtx& c, uint32_t n) {
  return ((n & 1u) << c.grp_pos[0]) | (((n >> 1) & 1u) << c.grp_pos[1]) | (((n >> 2) & 1u) << c.grp_pos[2]);
}

// Per-lane swizzled element offsets (in double2 units) relative to a batch: loads Pl[v] (B fragment register v,
// k = q + 4v, column n = g) and stores Ps[i] (D register i, row m = g + 8(i>>1), column n = 2q + (i&1)).
QCB_HD void dmma_lane_setup(const DmmaCtx& c, uint32_t lane, uint32_t (&Pl)[4], uint32_t (&Ps)[4], uint32_t& comp_l, uint32_t& comp_s) {
  const uint32_t g = lane >> 2, q = lane & 3u;
  comp_l = q & 1u;
  comp_s = g & 1u;
#pragma unroll
  for (uint32_t v = 0; v < 4; ++v) Pl[v] = swz(dmma_lane_offset(c, g) | dmma_pattern_offset(c, q + 4u * v, c.j_load), c.c);
#pragma unroll
  for (uint32_t i = 0; i < 4; ++i) Ps[i] = swz(dmma_lane_offset(c, 2u * q + (i & 1u)) | dmma_pattern_offset(c, g + 8u * (i >> 1), c.j_store), c.c);
}

// batch index -> tile-local base offset (group bits 3.. deposited at grp_pos[3..])
QCB_HD uint32_t dmma_batch_base(const DmmaCtx& c, uint32_t batch) {
  uint32_t o = 0;
#pragma unroll
  for (uint32_t i = 0; i < 7; ++i) if (i + 3u < c.n_grp) o |= ((batch >> i) & 1u) << c.grp_pos[i + 3];
  return o;
}

// variant index of a batch: bit j = value of condition bit cond_pos[j] (tile-local -> from base, else from ext_hi)
QCB_HD uint32_t dmma_variant(const DmmaCtx& c, uint32_t base, uint64_t ext_hi, uint32_t m) {
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    if (j < c.k) {
      const uint32_t p = c.cond_pos[j];
      const uint32_t bit = (p < m) ? ((base >> p) & 1u) : (uint32_t)((ext_hi >> (p - m)) & 1ULL);
      v |= bit << j;
    }
  }
  return v;
}

// ---- precomputed tables of a tensor-core round (built once per launch, shared by every tile of the stage).
// lane entry: 8 byte offsets inside the tile buffer: [0..4) loads (B register v), [4..8) stores (D register i); the
// re/im component select is folded in as bit 3.  XOR with the batch's swizzled byte offset gives the address
// (swz is linear over XOR and batch / lane / pattern bits are disjoint).
QCB_HD void dmma_lane_entry(const DmmaCtx& c, uint32_t lane, uint32_t (&e)[8]) {
  uint32_t Pl[4], Ps[4], cl, cs;
  dmma_lane_setup(c, lane, Pl, Ps, cl, cs);
#pragma unroll
  for (int v = 0; v < 4; ++v) e[v] = (Pl[v] << 4) | (cl << 3);
#pragma unroll
  for (int i = 0; i < 4; ++i) e[4 + i] = (Ps[i] << 4) | (cs << 3);
}

// batch entry: swizzled byte offset of the batch base (low 20 bits) | variant bits contributed by tile-local
// condition bits (<< 20).  The remaining variant bits come from the tile id / rank (dmma_variant_hi).
constexpr uint32_t DMMA_BATCH_OFF_MASK = 0xfffffu;
QCB_HD uint32_t dmma_batch_entry(const DmmaCtx& c, uint32_t batch, uint32_t m) {
  const uint32_t base = dmma_batch_base(c, batch);
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    if (j < c.k && c.cond_pos[j] < m) v |= ((base >> c.cond_pos[j]) & 1u) << j;
  return (swz(base, c.c) << 4) | (v << 20);
}

QCB_HD uint32_t dmma_variant_hi(const DmmaCtx& c, uint64_t ext_hi, uint32_t m) {
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    if (j < c.k && c.cond_pos[j] >= m) v |= (uint32_t)((ext_hi >> (c.cond_pos[j] - m)) & 1ULL) << j;
  return v;
}

// ---- "k3" tensor-core rounds (round kind 2): the dense 8x8 complex block M = Mr + i Mi of a round applied with Gauss's
// three real products instead of four.  With the amplitudes of a group as B = Br + i Bi:
//     K  = (Mr + Mi) Br            Re = K - Mi (Br + Bi)            Im = K + Mr (Bi - Br)
// i.e. three 8x8 real matrices P = Mr + Mi, N = -Mi, R = Mr (built by the host) and six m8n8k4 steps per 8 groups instead of
// the eight of the 16x16 real form; K is the C operand of both remaining products, so no extra additions on the result side.
// Fragment roles (PTX ISA mma.m8n8k4 .f64): A reg: row lane/4, col lane%4 (+4 per k-step); B reg: row lane%4 (+4 per k-step),
// col lane/4; C/D regs {0,1}: row lane/4, cols 2(lane%4) + {0,1}.  A lane therefore loads the COMPLEX amplitudes
// (pattern k = lane%4 + 4s, group lane/4), s = 0, 1, with two 16-byte loads and stores the complex results (pattern lane/4,
// groups 2(lane%4) + {0,1}) with two 16-byte stores.
// Round words [34] / [35] hold the slot roles: kmap = slot index carried by k-index bit 0 | bit 1 << 4 | bit 2 << 8 (columns of
// the matrices = input patterns), mmap likewise for the m-index (rows = output patterns).
struct K3Ctx {
  uint32_t n_grp, k, c;
  uint32_t slot_pos[3], grp_pos[10], cond_pos[4];
  uint32_t kmap[3], mmap[3];
  uint32_t mmap2[3];                         // round kind 3 only (see "paired rounds" below)
  uint64_t mat_off;
  uint32_t far_off[4], far_n[4];             // far-phase tables (word offset from the stage, entries): after block 1, after block 2,
                                             // before block 1, before block 2
};
constexpr uint32_t K3_FRAG_DOUBLES = 192;   // per variant: 6 A registers x 32 lanes (P0 P1 N0 N1 R0 R1)
constexpr uint32_t K3X_FRAG_DOUBLES = 384;  // round kind 3: the six registers of the first block, then those of the second

QCB_HD void decode_k3(const uint64_t* stage, uint32_t round_idx, K3Ctx& c) {
  const uint64_t* w = stage + T_STAGE_WORDS + (uint64_t)round_idx * T_ROUND_WORDS;
  c.c = (uint32_t)stage[43];
  c.mat_off = w[2]; c.n_grp = (uint32_t)w[18]; c.k = (uint32_t)w[29];
  for (int j = 0; j < 3; ++j) {
    c.slot_pos[j] = (uint32_t)w[4 + j];
    c.kmap[j] = (uint32_t)(w[34] >> (4 * j)) & 15u;
    c.mmap[j] = (uint32_t)(w[35] >> (4 * j)) & 15u;
    c.mmap2[j] = (uint32_t)(w[36] >> (4 * j)) & 15u;
  }
  for (int j = 0; j < 10; ++j) c.grp_pos[j] = (uint32_t)w[19 + j];
  for (int j = 0; j < 4; ++j) c.cond_pos[j] = (uint32_t)w[30 + j];
  c.far_off[0] = (uint32_t)w[37]; c.far_off[1] = (uint32_t)w[38]; c.far_off[2] = (uint32_t)(w[37] >> 32); c.far_off[3] = (uint32_t)(w[38] >> 32);
  for (int t = 0; t < 4; ++t) c.far_n[t] = (uint32_t)(w[39] >> (8 * t)) & 0xffu;
}
// ---- far phases (plan.cpp: classify_far / build_far_table).  A table entry is five words: position of a far bit (relative to
// the tile bits, i.e. a bit of ext_hi), gamma, phi_0, phi_1, phi_2.  For the tile ext_hi the block's rows are scaled by
//   d(pattern) = exp(i (G + sum_j (pattern bit j ? +F_j : -F_j))),   G / F_j = sums of gamma / phi_j over the entries whose bit is set.
// A table applied BEFORE the block scales the COLUMNS instead (input pattern; hardware k-index through kmap, or - second block of a
// pair - the lane-group index itself).
// far_sums: the four sums (sequential - the kernel does the same with a warp reduction over the entries).
QCB_HD void far_sums(const uint64_t* tab, uint32_t n, uint64_t ext_hi, double (&s)[4]) {
  s[0] = s[1] = s[2] = s[3] = 0.0;
  for (uint32_t e = 0; e < n; ++e) {
    if (!((ext_hi >> tab[5 * e]) & 1ULL)) continue;
    for (int k = 0; k < 4; ++k) s[k] += as_double(tab[5 * e + 1 + k]);
  }
}
QCB_HD double far_angle(const double (&s)[4], uint32_t pattern) {
  return s[0] + ((pattern & 1u) ? s[1] : -s[1]) + ((pattern & 2u) ? s[2] : -s[2]) + ((pattern & 4u) ? s[3] : -s[3]);
}
// slot pattern (slot-order bits) of a hardware m-index under a map (mmap / mmap2)
QCB_HD uint32_t k3_pattern_index(uint32_t idx, const uint32_t (&map)[3]) {
  return (((idx >> 0) & 1u) << map[0]) | (((idx >> 1) & 1u) << map[1]) | (((idx >> 2) & 1u) << map[2]);
}
// row scaling of the fragments of one lane: (P, N, R) = (Mr + Mi, -Mi, Mr) of a row multiplied by d = dr + i di
QCB_HD void far_scale(double& P, double& N, double& R, double dr, double di) {
  const double r2 = dr * R + di * N, n2 = dr * N - di * R;
  R = r2; N = n2; P = r2 - n2;
}
// tile-local offset of the slot pattern selected by a k-index / m-index
QCB_HD uint32_t k3_pattern_offset(const K3Ctx& c, uint32_t idx, const uint32_t (&map)[3]) {
  uint32_t o = 0;
#pragma unroll
  for (uint32_t b = 0; b < 3; ++b) {
    const uint32_t j = map[b];
    const uint32_t pos = (j == 0) ? c.slot_pos[0] : (j == 1 ? c.slot_pos[1] : c.slot_pos[2]);
    o |= ((idx >> b) & 1u) << pos;
  }
  return o;
}
QCB_HD uint32_t k3_group_offset(const K3Ctx& c, uint32_t n) {
  return ((n & 1u) << c.grp_pos[0]) | (((n >> 1) & 1u) << c.grp_pos[1]) | (((n >> 2) & 1u) << c.grp_pos[2]);
}
// lane entry: byte offsets inside the tile buffer of the two loads (k-steps 0, 1) and the two stores (result columns 0, 1)
QCB_HD void k3_lane_entry(const K3Ctx& c, uint32_t lane, uint32_t (&e)[4]) {
  const uint32_t g = lane >> 2, q = lane & 3u;
#pragma unroll
  for (uint32_t s = 0; s < 2; ++s) e[s] = swz(k3_group_offset(c, g) | k3_pattern_offset(c, q + 4u * s, c.kmap), c.c) << 4;
#pragma unroll
  for (uint32_t i = 0; i < 2; ++i) e[2 + i] = swz(k3_group_offset(c, 2u * q + i) | k3_pattern_offset(c, g, c.mmap), c.c) << 4;
}
// ---- paired rounds (round kind 3): TWO dense 8x8 complex blocks, on disjoint slot triples S1 and S2, in ONE pass over the
// shared tile.  The m8n8k4 fragment layouts transpose for free: after the first block a lane holds, as D registers i = 0, 1,
// the amplitudes (S1 pattern lane/4, group 2(lane%4) + i); handed to the next product as B registers of k-steps 0, 1 the
// hardware reads them as (k = lane%4 + 4i, column lane/4) - i.e. the batch's three lane-group bits have become the
// contraction index and the first block's output pattern has become the column.  So when S2 = the lane bits grp_pos[0..2] of
// the first block, the second block needs no trip through shared memory, no shuffle and no move: its matrix columns are
// simply stored in the order the hardware enumerates them (hardware k <-> group index 2(k & 3) + (k >> 2)).
// Loads are those of a kind-2 round on S1 (kmap).  After the second block a lane holds (S2 pattern lane/4 via mmap2,
// S1 pattern 2(lane%4) + i via mmap): its two stores.  Round word [36] = mmap2: index (0..2) into grp_pos[0..2] of the S2 bit
// carried by bit b of the second block's m-index, 4 bits each.
QCB_HD uint32_t k3x_hw_k_to_group(uint32_t k) { return 2u * (k & 3u) + (k >> 2); }
QCB_HD uint32_t k3x_pattern2_offset(const K3Ctx& c, uint32_t idx) {
  uint32_t o = 0;
#pragma unroll
  for (uint32_t b = 0; b < 3; ++b) {
    const uint32_t j = c.mmap2[b];
    const uint32_t pos = (j == 0) ? c.grp_pos[0] : (j == 1 ? c.grp_pos[1] : c.grp_pos[2]);
    o |= ((idx >> b) & 1u) << pos;
  }
  return o;
}
QCB_HD void k3x_lane_entry(const K3Ctx& c, uint32_t lane, uint32_t (&e)[4]) {
  const uint32_t g = lane >> 2, q = lane & 3u;
#pragma unroll
  for (uint32_t s = 0; s < 2; ++s) e[s] = swz(k3_group_offset(c, g) | k3_pattern_offset(c, q + 4u * s, c.kmap), c.c) << 4;
#pragma unroll
  for (uint32_t i = 0; i < 2; ++i) e[2 + i] = swz(k3x_pattern2_offset(c, g) | k3_pattern_offset(c, 2u * q + i, c.mmap), c.c) << 4;
}
// tile-local (unswizzled) index of the amplitude a lane stores as result column i of a batch with base 0: where the last
// round of a sweep writes it in global memory (direct store, stage flag T_FLAG_DIRECT_STORE)
QCB_HD uint32_t k3_lane_store_index(const K3Ctx& c, uint32_t lane, uint32_t i) {
  const uint32_t g = lane >> 2, q = lane & 3u;
  return k3_group_offset(c, 2u * q + i) | k3_pattern_offset(c, g, c.mmap);
}
QCB_HD uint32_t k3_batch_base(const K3Ctx& c, uint32_t batch) {
  uint32_t o = 0;
#pragma unroll
  for (uint32_t i = 0; i < 7; ++i) if (i + 3u < c.n_grp) o |= ((batch >> i) & 1u) << c.grp_pos[i + 3];
  return o;
}
// batch entry: swizzled byte offset of the batch base (low 20 bits) | variant bits of tile-local condition bits << 20
QCB_HD uint32_t k3_batch_entry(const K3Ctx& c, uint32_t batch, uint32_t m) {
  const uint32_t base = k3_batch_base(c, batch);
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    if (j < c.k && c.cond_pos[j] < m) v |= ((base >> c.cond_pos[j]) & 1u) << j;
  return (swz(base, c.c) << 4) | (v << 20);
}
QCB_HD uint32_t k3_variant_hi(const K3Ctx& c, uint64_t ext_hi, uint32_t m) {
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    if (j < c.k && c.cond_pos[j] >= m) v |= (uint32_t)((ext_hi >> (c.cond_pos[j] - m)) & 1ULL) << j;
  return v;
}

// ---- tile addressing
struct StageCtx {
  uint32_t n_local, m, L, n_rounds, n_runs, c;
  uint64_t ext_hi_base, skip_mask, skip_val;
};

QCB_HD void decode_stage(const uint64_t* st, StageCtx& sc) {
  sc.n_local = (uint32_t)st[0]; sc.m = (uint32_t)st[1]; sc.L = (uint32_t)st[2]; sc.n_rounds = (uint32_t)st[3];
  sc.n_runs = (uint32_t)st[4]; sc.ext_hi_base = st[5]; sc.skip_mask = st[6]; sc.skip_val = st[7]; sc.c = (uint32_t)st[43];
}

// active-tile ordinal -> tile id (deposit into the tile-id bits not fixed by skip_mask, OR the fixed value)
QCB_HD uint64_t active_to_tile(const StageCtx& sc, uint64_t a) {
  const uint32_t nb = sc.n_local - sc.m;
  const uint64_t tmask = (nb >= 64) ? ~0ULL : ((1ULL << nb) - 1ULL);
  const uint64_t fixed = sc.skip_mask & tmask;
  if (!fixed) return a;
  uint64_t t = sc.skip_val & fixed, src = a;
  for (uint32_t b = 0; b < nb; ++b)
    if (!((fixed >> b) & 1ULL)) { t |= (src & 1ULL) << b; src >>= 1; }
  return t;
}

// tile id -> base amplitude offset (tile bits = 0): scatter the id over the runs of non-tile positions
QCB_HD uint64_t tile_base(const uint64_t* st, const StageCtx& sc, uint64_t tile) {
  uint64_t base = 0, t = tile;
  for (uint32_t k = 0; k < sc.n_runs; ++k) {
    const uint32_t start = (uint32_t)(st[24 + k] & 0xff), len = (uint32_t)(st[24 + k] >> 8);
    base |= (t & ((1ULL << len) - 1ULL)) << start;
    t >>= len;
  }
  return base;
}

// offset contributed by the high tile bits (tile-local bits L..m-1) for hi = local >> L
QCB_HD uint64_t hi_offset_from(const uint64_t* st, const StageCtx& sc, uint32_t hi, uint32_t from) {
  uint64_t o = 0;
  for (uint32_t k = from; k < sc.m; ++k) o |= (uint64_t)((hi >> (k - from)) & 1u) << st[8 + k];
  return o;
}
QCB_HD uint64_t hi_offset(const uint64_t* st, const StageCtx& sc, uint32_t hi) { return hi_offset_from(st, sc, hi, sc.L); }

}  // namespace qcb
