// kernels.cu — hand-written sm_100a kernels of libqcb200.so.
//
//  k_tile_stage        the fused gate executor: one launch = one sweep over the local state.  One persistent,
//                      warp-specialised CTA per SM: mover warps stream tiles of 2^m amplitudes through a ring
//                      of shared-memory buffers (cp.async / st.global, or TMA tensor copies), two consumer
//                      groups apply the stage's rounds - dense 8x8 complex blocks on the fp64 tensor cores
//                      (DMMA.8x8x4) or interpreter rounds (tile_core.h) - and the tile is written back in
//                      place.  Algorithmic bytes = 32 * 2^n_local * (fraction of tiles visited); HBM-bound up
//                      to two rounds per sweep, fp64-tensor-bound beyond.
//  reductions          norm^2 / complex sum / Pauli expectation / marginal histogram: streaming reads
//                      (16 B per amplitude), warp-shuffle + shared-memory block reduce, deterministic
//                      two-pass finalisation (no floating-point atomics).
//  sampling            chunk sums -> scan of chunk sums -> per-shot binary search + in-chunk scan;
//                      implements the reference's measure-state rule (domain/state.clj:894-913).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "kernels.h"
#include "tile_core.h"

namespace qcb {

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of K values per thread; result valid in thread 0.  smem: K * 32 doubles
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sm) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sm[k * 32 + wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = (lane < nw) ? sm[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ mbarrier / named-barrier primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
// BACKOFF: the waiter has nothing else to do for a long time (the mover waiting for a tile to be finished): let the hardware
// suspend the thread inside try_wait (suspend-time hint, ns) instead of polling - a polling warp competes for issue slots and
// for the shared-memory pipe with the consumers (round 1: 27 % of all stall samples sat on the two polling branches).
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  for (;;) {
    if (BACKOFF)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(a), "r"(parity), "r"(20000u) : "memory");
    else
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) break;
  }
}
// arrive on `bar` once every cp.async issued so far by this thread has landed in shared memory
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// the round barrier of consumer group `grp`: barrier 1 + grp with the id as an IMMEDIATE (a barrier id held in a register
// is legal PTX, but compute-sanitizer's racecheck does not follow it and then reports every round-to-round hand-over)
template <uint32_t NT>
__device__ __forceinline__ void group_bar_sync(uint32_t grp) {
  switch (grp) {
    case 0: asm volatile("bar.sync 1, %0;\n" ::"n"(NT) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;\n" ::"n"(NT) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;\n" ::"n"(NT) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;\n" ::"n"(NT) : "memory"); break;
  }
}

// ------------------------------------------------------------------ TMA (cp.async.bulk.tensor) primitives
// The state is described to the TMA unit as a 2-D tensor of doubles [rows = 2^(n_local-3)][16] (one row = 8 amplitudes =
// 128 bytes) with the 128-byte swizzle; a box is one contiguous run of 2^c amplitudes = 2^(c-3) rows.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_rows(uint32_t dst_s, const CUtensorMap* map, int32_t row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(dst_s), "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_rows(const CUtensorMap* map, int32_t row, uint32_t src_s) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(row), "r"(src_s) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// make this thread's generic-proxy writes to shared memory visible to the async proxy (TMA store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------ optional cycle accounting (make PROFILE=1)
#ifdef QCB_TILE_PROFILE
// ablation switches of the profiling build (QCB_TILE_DBG): 1 = skip the DMMAs, 2 = skip the fragment LDS/STS,
// 4 = skip the HBM traffic of the mover, 8 = skip the round barriers, 16 = fetch the A fragments once only
// (results are then wrong on purpose: timing experiments only)
__device__ int g_tile_dbg;
#ifdef QCB_TILE_ABLATE
__device__ __forceinline__ int tile_dbg() { int v; asm volatile("ld.global.cv.s32 %0, [%1];" : "=r"(v) : "l"(&g_tile_dbg)); return v; }
#define DBG_DECL const int dbg_bits = tile_dbg()
#define DBG_ON(bit) (dbg_bits & (bit))
#else
#define DBG_DECL const int dbg_bits = 0; (void)dbg_bits
#define DBG_ON(bit) false
#endif
enum { PF_C_WAIT_FULL = 0, PF_C_BARRIER, PF_C_SETUP, PF_C_ROUND, PF_C_TOTAL, PF_M_LOAD, PF_M_WAIT_DONE, PF_M_STORE, PF_M_TOTAL, PF_N };
__device__ unsigned long long g_tile_prof[PF_N];
// per-warp accumulation in registers, one atomic per category at the end of the kernel
#define PF_DECL long long pf_t = clock64(), pf_t0 = pf_t; unsigned long long pf_acc[PF_N] = {}
#define PF_ADD(cat) do { const long long pf_n = clock64(); pf_acc[cat] += (unsigned long long)(pf_n - pf_t); pf_t = pf_n; } while (0)
#define PF_TOTAL(cat) do { pf_acc[cat] = (unsigned long long)(clock64() - pf_t0); if ((threadIdx.x & 31) == 0) { for (int pf_i = 0; pf_i < PF_N; ++pf_i) if (pf_acc[pf_i]) atomicAdd(&g_tile_prof[pf_i], pf_acc[pf_i]); } } while (0)
void tile_prof_dump() {
  unsigned long long h[PF_N];
  if (cudaMemcpyFromSymbol(h, g_tile_prof, sizeof h) != cudaSuccess) return;
  static const char* names[PF_N] = {"consumer wait full", "consumer round barrier", "consumer setup+prefetch", "consumer round work", "consumer total",
                                    "mover load issue", "mover wait done", "mover store", "mover total"};
  for (int i = 0; i < PF_N; ++i) {
    const unsigned long long tot = h[i < PF_M_LOAD ? PF_C_TOTAL : PF_M_TOTAL];
    fprintf(stderr, "[tile-prof] %-26s %14llu warp-cycles  %5.1f%%\n", names[i], h[i], tot ? 100.0 * (double)h[i] / (double)tot : 0.0);
  }
  unsigned long long z[PF_N] = {};
  cudaMemcpyToSymbol(g_tile_prof, z, sizeof z);
}
#else
#define DBG_DECL const int dbg_bits = 0; (void)dbg_bits
#define DBG_ON(bit) false
#define PF_DECL
#define PF_ADD(cat)
#define PF_TOTAL(cat)
void tile_prof_dump() {}
#endif

// ------------------------------------------------------------------ optional timeline trace (-DQCB_TILE_TRACE)
// Time stamps of one CTA's warps at the hand-over points of the tile pipeline (tile arrived, pass started, first operands in
// flight, steady-state loop entered / left, results stored, tile released; mover: loads issued, tile written back), for a few
// tiles of every launch.  Dumped as text when the handle is destroyed; scripts/trace_summary.py turns it into per-segment
// cycle counts.  Timing experiments only.
#ifdef QCB_TILE_TRACE
__device__ unsigned long long g_tile_trace[1 << 16];
__device__ unsigned g_tile_trace_n;
struct TileTrace {
  uint32_t warp, j, r;
  bool on;
  __device__ __forceinline__ void operator()(uint32_t id) const {
    if (!on) return;
    const unsigned k = atomicAdd(&g_tile_trace_n, 1u);
    if (k < (1u << 16))
      g_tile_trace[k] = ((unsigned long long)id << 56) | ((unsigned long long)warp << 52) | ((unsigned long long)(j & 0xffu) << 44) |
                        ((unsigned long long)(r & 0xfu) << 40) | ((unsigned long long)clock64() & 0xffffffffffULL);
  }
};
#define TRACE_MAKE(w, jj, rr) TileTrace{(w), (jj), (rr), blockIdx.x == 3u && (threadIdx.x & 31u) == 0u && (jj) >= 40u && (jj) < 44u}
void tile_trace_dump() {
  unsigned n = 0;
  if (cudaMemcpyFromSymbol(&n, g_tile_trace_n, sizeof n) != cudaSuccess || n == 0) return;
  if (n > (1u << 16)) n = 1u << 16;
  unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * n);
  if (cudaMemcpyFromSymbol(h, g_tile_trace, sizeof(unsigned long long) * n) == cudaSuccess)
    for (unsigned i = 0; i < n; ++i)
      fprintf(stderr, "[tile-trace] %u %u %u %u %llu\n", (unsigned)(h[i] >> 56), (unsigned)(h[i] >> 52) & 15u, (unsigned)(h[i] >> 44) & 255u,
              (unsigned)(h[i] >> 40) & 15u, h[i] & 0xffffffffffULL);
  free(h);
  unsigned z = 0;
  cudaMemcpyToSymbol(g_tile_trace_n, &z, sizeof z);
}
#else
struct TileTrace { __device__ __forceinline__ void operator()(uint32_t) const {} };
#define TRACE_MAKE(w, jj, rr) TileTrace{}
void tile_trace_dump() {}
#endif

// ------------------------------------------------------------------ tensor-core round
// One warp applies the round's dense 16x16 real matrix to 8 groups at a time:
//   D(16x8) = A(16x16) * B(16x8),  A = matrix variant (fragments from global through L1, reloaded only when the
//   variant changes), B column n = the 16 reals of group n of the batch, gathered from the swizzled shared tile with
//   8-byte loads in fragment order, D scattered back in place.  Fragment layouts: PTX ISA mma.m16n8k16 .f64
//   (A reg i: row lane/4 + 8(i&1), col lane%4 + 4(i>>1); B reg v: row lane%4 + 4v, col lane/4;
//    D reg i: row lane/4 + 8(i>>1), col 2(lane%4) + (i&1)).
__device__ __forceinline__ void dmma_m16n8k16(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
      "{%16,%17,%18,%19};\n"
      : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]), "d"(0.0), "d"(0.0), "d"(0.0), "d"(0.0));
}

// One k-step of the 16x8x16 product as the native 8x8x4 operation (the m16n8k16 fragments ARE the m8n8k4 fragments:
// A reg i = m-half (i & 1), k-step (i >> 1); B reg v = k-step v; D regs {0,1} / {2,3} = m-half 0 / 1).  Issuing the eight
// steps as separate instructions lets the round interleave its shared-memory traffic between them (dmma_round_run).
__device__ __forceinline__ void dmma_884_first(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}
__device__ __forceinline__ void dmma_884_acc(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Per-round tables live in shared memory (built once per launch): lane_tab[r][lane] = 8 byte offsets
// (tile_core.h: dmma_lane_entry), batch_tab[r][b] = swizzled byte offset | local variant bits << 20.
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dmma_load_A(double (&A)[8], const double* __restrict__ mats, uint32_t var) {
#pragma unroll
  for (int i = 0; i < 8; ++i) A[i] = __ldg(mats + ((size_t)var * 8 + i) * 32);
}

// Variant bits contributed by the tile-id / rank condition bits of a round.  hi_desc packs, for each of the (at most 4)
// condition bits, a 6-bit field: the bit's position above the tile bits, or 63 when the bit is tile-local / unused.
__device__ __forceinline__ uint32_t dmma_var_hi(uint32_t hi_desc, uint64_t ext_hi) {
  uint32_t v = 0;
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    const uint32_t f = (hi_desc >> (6u * j)) & 63u;
    if (f != 63u) v |= (uint32_t)((ext_hi >> f) & 1ULL) << j;
  }
  return v;
}

// One warp's share of a tensor-core round on the tile at shared address `tile_s`: batches btab[0..per).  A holds variant
// `cur` on entry.  Fast path (every batch of the warp uses the same variant): software pipeline
// LDS(i+1) | DMMA(i) | STS(i-1), so the warp's stream of tensor instructions is not interrupted by shared-memory latency.
__device__ __forceinline__ void dmma_round_run(uint32_t tile_s, const uint4* lane_tab_r, const uint32_t* btab, uint32_t per,
                                               uint32_t var_hi, const double* __restrict__ mats, uint32_t lane, double (&A)[8],
                                               uint32_t cur, int dbg_bits) {
  (void)dbg_bits;
  const uint4 l0 = lane_tab_r[2u * lane], l1 = lane_tab_r[2u * lane + 1u];
  const uint32_t pl[4] = {l0.x, l0.y, l0.z, l0.w};
  const uint32_t ps[4] = {l1.x, l1.y, l1.z, l1.w};
  auto load_B = [&](uint32_t X, double (&B)[4]) {
    if (DBG_ON(2)) { for (int v = 0; v < 4; ++v) B[v] = (double)(X + v); return; }
#pragma unroll
    for (int v = 0; v < 4; ++v) B[v] = lds_f64(tile_s + (pl[v] ^ X));
  };
  auto store_D = [&](uint32_t X, const double (&D)[4]) {
    if (DBG_ON(2)) { if (D[0] + D[1] + D[2] + D[3] == 1.2345) sts_f64(tile_s, D[0]); return; }
    __syncwarp();     // intra-warp in-place update: see k3_pp
#pragma unroll
    for (int i = 0; i < 4; ++i) sts_f64(tile_s + (ps[i] ^ X), D[i]);
  };
  const uint32_t e_first = btab[0], e_last = btab[per - 1u];
  if ((e_first >> 20) == (e_last >> 20)) {
    // local condition bits are the top bits of the batch index: equal at both ends => equal throughout
    // Each 8x8x4 step occupies the tensor pipe of the SM partition for 16 cycles and a warp issues in order: the loads of
    // batch i+1 and the stores of batch i-1 are slotted between the steps of batch i, so the pipe never waits for a warp
    // that is busy with its shared-memory traffic (two warps interleaving their steps otherwise finish - and idle - together).
    double Bc[4], Dp[4];
    uint32_t X = e_first & DMMA_BATCH_OFF_MASK, Xp = X;
    load_B(X, Bc);
    auto step = [&](uint32_t i, bool more) {
      double Bn[4], D[4];
      const uint32_t Xn = more ? (btab[i + 1u] & DMMA_BATCH_OFF_MASK) : X;
      if (DBG_ON(1)) {
        for (int q = 0; q < 4; ++q) D[q] = Bc[q];
        if (more) load_B(Xn, Bn);
        if (i) store_D(Xp, Dp);
      } else {
        dmma_884_first(D[0], D[1], A[0], Bc[0]);
        if (more) Bn[0] = lds_f64(tile_s + (pl[0] ^ Xn));
        dmma_884_first(D[2], D[3], A[1], Bc[0]);
        if (more) Bn[1] = lds_f64(tile_s + (pl[1] ^ Xn));
        dmma_884_acc(D[0], D[1], A[2], Bc[1]);
        if (more) Bn[2] = lds_f64(tile_s + (pl[2] ^ Xn));
        dmma_884_acc(D[2], D[3], A[3], Bc[1]);
        if (more) Bn[3] = lds_f64(tile_s + (pl[3] ^ Xn));
        dmma_884_acc(D[0], D[1], A[4], Bc[2]);
        if (i) __syncwarp();
        if (i) sts_f64(tile_s + (ps[0] ^ Xp), Dp[0]);
        dmma_884_acc(D[2], D[3], A[5], Bc[2]);
        if (i) sts_f64(tile_s + (ps[1] ^ Xp), Dp[1]);
        dmma_884_acc(D[0], D[1], A[6], Bc[3]);
        if (i) sts_f64(tile_s + (ps[2] ^ Xp), Dp[2]);
        dmma_884_acc(D[2], D[3], A[7], Bc[3]);
        if (i) sts_f64(tile_s + (ps[3] ^ Xp), Dp[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { Dp[q] = D[q]; Bc[q] = Bn[q]; }
      Xp = X; X = Xn;
    };
    // Each 8x8x4 step occupies the tensor pipe of the SM partition for 16 cycles and a warp issues in order: the loads of
    // batch i+1 and the stores of batch i-1 are slotted between the steps of batch i, so the pipe never waits for a warp
    // that is busy with its shared-memory traffic (two warps interleaving their steps otherwise finish - and idle - together).
    // The common trip counts are unrolled completely (no register rotation moves).
    if (per == 16u) {
#pragma unroll
      for (uint32_t i = 0; i < 16; ++i) step(i, i + 1u < 16u);
    } else if (per == 8u) {
#pragma unroll
      for (uint32_t i = 0; i < 8; ++i) step(i, i + 1u < 8u);
    } else {
#pragma unroll 2
      for (uint32_t i = 0; i < per; ++i) step(i, i + 1u < per);
    }
    store_D(Xp, Dp);
    return;
  }
  for (uint32_t bi = 0; bi < per; ++bi) {
    const uint32_t e = btab[bi];
    const uint32_t v = var_hi | (e >> 20), X = e & DMMA_BATCH_OFF_MASK;
    if (v != cur) { dmma_load_A(A, mats, v); cur = v; }
    double B[4], D[4];
    load_B(X, B);
    dmma_m16n8k16(D, A, B);
    store_D(X, D);
  }
}

// ------------------------------------------------------------------ tensor-core round, three-product form (round kind 2)
// tile_core.h: K3Ctx.  Per batch of 8 groups a lane issues two 16-byte loads (complex amplitudes of pattern lane%4 + 4s,
// group lane/4), four DADD (Br + Bi, Bi - Br), six DMMA.8x8x4 (K = P Br; Re = N (Br + Bi) + K; Im = R (Bi - Br) + K) and two
// 16-byte stores - against eight DMMA, four 8-byte loads and four 8-byte stores of the 16x16 real form.
__device__ __forceinline__ void lds_c128(uint32_t addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts_c128(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void dmma_884_c(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int NA>
__device__ __forceinline__ void k3_load_A(double (&A)[NA], const double* __restrict__ mats, uint32_t var) {
#pragma unroll
  for (int i = 0; i < 6; ++i) A[i] = __ldg(mats + (size_t)var * K3_FRAG_DOUBLES + i * 32);
}

// Ablation switches of the profiling build (make ABLATE=1, QCB_TILE_DBG): ABL & 1 = no tensor instructions, ABL & 2 = no
// fragment loads / stores (results are then wrong on purpose: timing experiments only).  ABL = 0 in the product build.
template <int ABL>
__device__ __forceinline__ void xmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  if (ABL & 1) { d0 = c0; d1 = c1; } else dmma_884_c(d0, d1, a, b, c0, c1);
}
template <int ABL>
__device__ __forceinline__ void xlds(uint32_t addr, double& x, double& y) {
  if (ABL & 2) { x = __hiloint2double(0x3ff00000 | (int)(addr & 0xfffffu), (int)addr); y = __hiloint2double(0x3fe00000 | (int)(addr & 0xfffffu), 7); } else lds_c128(addr, x, y);
}
template <int ABL>
__device__ __forceinline__ void xsts(uint32_t addr, double x, double y) {
  if (ABL & 2) { if (addr == 0xffffffffu) sts_c128(addr, x, y); } else sts_c128(addr, x, y);     // never taken, keeps the results alive
}

// the four sums of a far table for the tile ext_hi: lane e takes entry e (at most 32 entries), xor-shuffle reduction
__device__ __forceinline__ void far_sums_warp(const uint64_t* __restrict__ tab, uint32_t n, uint64_t ext_hi, uint32_t lane, double (&s)[4]) {
  s[0] = s[1] = s[2] = s[3] = 0.0;
  if (lane < n) {
    const uint64_t* e = tab + 5u * lane;
    if ((ext_hi >> __ldg(e)) & 1ULL) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] = __longlong_as_double((long long)__ldg(e + 1 + k));
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
}

// Where a round's results go.  Shared tile: byte address tile_s + (lane offset ^ batch offset).  DIRECT (the last round of a
// sweep, stage flag T_FLAG_DIRECT_STORE): straight to global memory from registers - amplitude offset (lane part ^ batch
// part) from the tile's base; the batch parts come from gtab (built once per launch), the lane parts are per-round constants.
struct K3Out {
  uint32_t tile_s;            // shared tile
  uint32_t lz, lw;            // lane byte offsets of result columns 0 / 1 in the shared tile
  double2* gbase;             // DIRECT: tile base in global memory
  uint64_t g0, g1;            // DIRECT: lane parts of the global amplitude offsets of result columns 0 / 1
  const uint64_t* gtab;       // DIRECT: batch parts, indexed like btab
};
template <bool DIRECT>
__device__ __forceinline__ void k3_store(const K3Out& o, uint32_t X, uint64_t G, double r0, double i0, double r1, double i1) {
  if (DIRECT) {
    __stcs(o.gbase + (o.g0 ^ G), double2{r0, i0});
    __stcs(o.gbase + (o.g1 ^ G), double2{r1, i1});
  } else {
    __syncwarp();     // intra-warp in-place update: see k3_pp
    sts_c128(o.tile_s + (o.lz ^ X), r0, i0);
    sts_c128(o.tile_s + (o.lw ^ X), r1, i1);
  }
}

// Batches btab[0..per) of one warp, all with the matrix variant held in A (P0 P1 N0 N1 R0 R1).  Software pipeline over
// batches: in the steady state the warp issues, for batch i,   Re0(i) K0(i+1) Im0(i) K1(i+1) Re1(i) Im1(i)   - every
// dependent pair is at least two tensor instructions apart (a DMMA.8x8x4 holds the pipe of the SM partition for 16 cycles,
// its result is ready after 26), so a single warp can keep the pipe busy - with the DADDs of batch i+1, the loads of batch
// i+2 and the stores of batch i-1 slotted in between.  The loop stays ROLLED on purpose: ptxas re-schedules a fully
// unrolled body batch by batch (dependent K0 -> K1 -> Re0 back to back, loads sunk to their first use), which undoes the
// pipeline; across a loop back edge it cannot.  (-DQCB_K3_UNROLL builds the unrolled variant for comparison.)
struct K3State {
  double K0, K1, s0, s1, d0, d1;              // batch i: K = P Br and (Br + Bi), (Bi - Br) of k-steps 0, 1
  double nr0, ni0, nr1, ni1;                  // batch i+1: raw loads
  double pr0, pr1, pi0, pi1;                  // batch i-1: results waiting for their stores
  uint32_t X, Xp, Xn;                         // swizzled byte offsets of batches i, i-1, i+1
  uint64_t Gc, Gp;                            // DIRECT: global batch offsets of batches i, i-1
};
template <bool HAS1, bool HAS2, bool HASP, bool DIRECT>
__device__ __forceinline__ void k3_iter(K3State& t, uint32_t tile_s, const uint4& lt, uint32_t x_next2, const double (&A)[12],
                                        const K3Out& out, uint64_t g_next) {
  double re0, re1, im0, im1, k0n = 0, k1n = 0, ns0 = 0, ns1 = 0, nd0 = 0, nd1 = 0;
  dmma_884_c(re0, re1, A[2], t.s0, t.K0, t.K1);
  if (HAS1) dmma_884_c(k0n, k1n, A[0], t.nr0, 0.0, 0.0);
  dmma_884_c(im0, im1, A[4], t.d0, t.K0, t.K1);
  if (HAS1) {
    dmma_884_c(k0n, k1n, A[1], t.nr1, k0n, k1n);
    ns0 = t.nr0 + t.ni0; nd0 = t.ni0 - t.nr0; ns1 = t.nr1 + t.ni1; nd1 = t.ni1 - t.nr1;
  }
  if (HAS2) {
    lds_c128(tile_s + (lt.x ^ x_next2), t.nr0, t.ni0);
    lds_c128(tile_s + (lt.y ^ x_next2), t.nr1, t.ni1);
  }
  dmma_884_c(re0, re1, A[3], t.s1, re0, re1);
  if (HASP) k3_store<DIRECT>(out, t.Xp, t.Gp, t.pr0, t.pi0, t.pr1, t.pi1);
  dmma_884_c(im0, im1, A[5], t.d1, im0, im1);
  t.pr0 = re0; t.pr1 = re1; t.pi0 = im0; t.pi1 = im1;
  t.Xp = t.X; t.X = t.Xn; t.Xn = x_next2;
  if (DIRECT) { t.Gp = t.Gc; t.Gc = g_next; }
  t.K0 = k0n; t.K1 = k1n; t.s0 = ns0; t.s1 = ns1; t.d0 = nd0; t.d1 = nd1;
}
template <bool DIRECT>
__device__ __forceinline__ void k3_batches(uint32_t tile_s, const uint4 lt, const uint32_t* btab, uint32_t per, const double (&A)[12],
                                           const K3Out& out) {
  K3State t;
  t.X = btab[0] & DMMA_BATCH_OFF_MASK; t.Xp = t.X; t.Xn = t.X;
  t.Gc = DIRECT ? out.gtab[0] : 0; t.Gp = t.Gc;
  // g_next of iteration i = global offset of batch i+1 (it becomes Gc when the state rotates)
  auto gn = [&](uint32_t i) -> uint64_t { return (DIRECT && i + 1u < per) ? out.gtab[i + 1u] : 0; };
  t.pr0 = t.pr1 = t.pi0 = t.pi1 = 0.0;
  {
    double r0, i0, r1, i1;
    lds_c128(tile_s + (lt.x ^ t.X), r0, i0);
    lds_c128(tile_s + (lt.y ^ t.X), r1, i1);
    if (per > 1u) {
      t.Xn = btab[1] & DMMA_BATCH_OFF_MASK;
      lds_c128(tile_s + (lt.x ^ t.Xn), t.nr0, t.ni0);
      lds_c128(tile_s + (lt.y ^ t.Xn), t.nr1, t.ni1);
    } else { t.nr0 = t.ni0 = t.nr1 = t.ni1 = 0.0; }
    t.s0 = r0 + i0; t.d0 = i0 - r0; t.s1 = r1 + i1; t.d1 = i1 - r1;
    dmma_884_c(t.K0, t.K1, A[0], r0, 0.0, 0.0);
    dmma_884_c(t.K0, t.K1, A[1], r1, t.K0, t.K1);
  }
  if (per >= 3u) {
    k3_iter<true, true, false, DIRECT>(t, tile_s, lt, btab[2] & DMMA_BATCH_OFF_MASK, A, out, gn(0));
#ifdef QCB_K3_UNROLL
#pragma unroll 16
#else
#pragma unroll 1
#endif
    for (uint32_t i = 1; i + 2u < per; ++i) k3_iter<true, true, true, DIRECT>(t, tile_s, lt, btab[i + 2u] & DMMA_BATCH_OFF_MASK, A, out, gn(i));
    k3_iter<true, false, true, DIRECT>(t, tile_s, lt, 0u, A, out, gn(per - 2u));
    k3_iter<false, false, true, DIRECT>(t, tile_s, lt, 0u, A, out, 0);
  } else if (per == 2u) {
    k3_iter<true, false, false, DIRECT>(t, tile_s, lt, 0u, A, out, gn(0));
    k3_iter<false, false, true, DIRECT>(t, tile_s, lt, 0u, A, out, 0);
  } else {
    k3_iter<false, false, false, DIRECT>(t, tile_s, lt, 0u, A, out, 0);
  }
  k3_store<DIRECT>(out, t.Xp, t.Gp, t.pr0, t.pi0, t.pr1, t.pi1);
}

// ---- the same pipeline without register rotation: two operand sets that swap roles every batch (the loop body handles two
// batches), so every value is consumed before the instruction that produces its successor and ptxas needs no moves.  Round 2
// profile of the rotating loop: 24 IMAD.MOV per batch, the ones that copy fresh DMMA results stall the warp on the short
// scoreboard, and the batch-offset LDS sits in the address chain of the loads.  Here the batch offset is fetched one batch
// ahead of its first use and the operand loads of batch i+2 are issued at the top of batch i.
struct K3Set {
  double K0, K1, s0, s1, d0, d1;     // K = P Br; Br + Bi and Bi - Br of k-steps 0, 1
  double r0, i0, r1, i1;             // raw loads
  uint32_t X;                        // swizzled byte offset of the batch the set currently belongs to
};
// one batch: c = its operands (K, s, d ready), n = the next batch's (raw loads in flight); HAS1 / HAS2 / HASP: a batch i+1 /
// i+2 / i-1 exists.  pr / pi = the results (stored at the top of the NEXT call), pa0 / pa1 their shared-memory addresses.
// DIRECT: pa0 / pa1 are unused, the results go to global memory at (lane part ^ pg), pg = the batch's global offset.
template <bool HAS1, bool HAS2, bool HASP, bool DIRECT, int ABL = 0>
__device__ __forceinline__ void k3_pp(K3Set& c, K3Set& n, double& pr0, double& pr1, double& pi0, double& pi1, uint32_t& pa0, uint32_t& pa1,
                                      uint64_t& pg, uint32_t tile_s, const uint4& lt, uint32_t xq, const double (&A)[12],
                                      const K3Out& out, uint64_t g_cur) {
  if (HASP) {
    // In-place update inside a warp: the amplitudes one lane stores were loaded (as B operands) by OTHER lanes of the same
    // warp.  Those loads fed mma.sync instructions that have completed before the results exist, so the order is given by
    // data dependence; the __syncwarp states it in the memory model's terms as well (and is what racecheck looks for).
    if (!DIRECT) __syncwarp();
    if (DIRECT) { __stcs(out.gbase + (out.g0 ^ pg), double2{pr0, pi0}); __stcs(out.gbase + (out.g1 ^ pg), double2{pr1, pi1}); }
    else { xsts<ABL>(pa0, pr0, pi0); xsts<ABL>(pa1, pr1, pi1); }
  }
  if (DIRECT) pg = g_cur;
  else { pa0 = tile_s + (lt.z ^ c.X); pa1 = tile_s + (lt.w ^ c.X); }
  if (HAS2) {
    c.X = xq & DMMA_BATCH_OFF_MASK;
    xlds<ABL>(tile_s + (lt.x ^ c.X), c.r0, c.i0);
    xlds<ABL>(tile_s + (lt.y ^ c.X), c.r1, c.i1);
  }
  xmma<ABL>(pr0, pr1, A[2], c.s0, c.K0, c.K1);
  if (HAS1) xmma<ABL>(n.K0, n.K1, A[0], n.r0, 0.0, 0.0);
  xmma<ABL>(pi0, pi1, A[4], c.d0, c.K0, c.K1);
  if (HAS1) {
    xmma<ABL>(n.K0, n.K1, A[1], n.r1, n.K0, n.K1);
    n.s0 = n.r0 + n.i0; n.d0 = n.i0 - n.r0;
  }
  xmma<ABL>(pr0, pr1, A[3], c.s1, pr0, pr1);
  if (HAS1) { n.s1 = n.r1 + n.i1; n.d1 = n.i1 - n.r1; }
  xmma<ABL>(pi0, pi1, A[5], c.d1, pi0, pi1);
}
// per even and >= 4
// (Also fetching the NEXT pass's lane entry and first batch entries here, so that a pass starts with its operand loads instead
// of dependent table look-ups, was measured slower: 5604 - 5639 vs 5716 - 5723 gates/s, profiles/r2u_ab.log.)
// mid(): work for the NEXT pass (operand prefetch) placed between the peeled calls, where its latency chain (table lookups,
// address arithmetic, fragment loads) hides behind this pass's tensor instructions instead of sitting between two passes
template <bool DIRECT, int ABL = 0, class F>
__device__ __forceinline__ void k3_batches_pp(uint32_t tile_s, const uint4 lt, const uint32_t* btab, uint32_t per, const double (&A)[12],
                                              const K3Out& out, F&& mid, const TileTrace tr = TileTrace{}) {
  K3Set a, b;
  double pr0 = 0, pr1 = 0, pi0 = 0, pi1 = 0;
  uint32_t pa0 = 0, pa1 = 0;
  uint64_t pg = 0;
  auto gq = [&](uint32_t i) -> uint64_t { return DIRECT ? out.gtab[i] : 0; };
  a.X = btab[0] & DMMA_BATCH_OFF_MASK;
  b.X = btab[1] & DMMA_BATCH_OFF_MASK;
  xlds<ABL>(tile_s + (lt.x ^ a.X), a.r0, a.i0);
  xlds<ABL>(tile_s + (lt.y ^ a.X), a.r1, a.i1);
  xlds<ABL>(tile_s + (lt.x ^ b.X), b.r0, b.i0);
  xlds<ABL>(tile_s + (lt.y ^ b.X), b.r1, b.i1);
  uint32_t xq = btab[2];
  tr(4);
  a.s0 = a.r0 + a.i0; a.d0 = a.i0 - a.r0; a.s1 = a.r1 + a.i1; a.d1 = a.i1 - a.r1;
  xmma<ABL>(a.K0, a.K1, A[0], a.r0, 0.0, 0.0);
  xmma<ABL>(a.K0, a.K1, A[1], a.r1, a.K0, a.K1);
  // batches 0, 1
  k3_pp<true, true, false, DIRECT, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, xq, A, out, gq(0));
  xq = btab[3];
  k3_pp<true, true, true, DIRECT, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, xq, A, out, gq(1));
  tr(5);
  mid();
  tr(6);
  // batches 2 .. per-3 (both look-aheads exist)
  // two loop bodies (four batches) per back edge: +2 % over one (profiles/r2g_ab.log); fully unrolled, ptxas serialises the batches
#ifdef QCB_PP_UNROLL1
#pragma unroll 1
#else
#pragma unroll 2
#endif
  for (uint32_t i = 2; i + 2u < per; i += 2u) {
    xq = btab[i + 2u];
    k3_pp<true, true, true, DIRECT, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, xq, A, out, gq(i));
    xq = btab[i + 3u];
    k3_pp<true, true, true, DIRECT, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, xq, A, out, gq(i + 1u));
  }
  tr(7);
  k3_pp<true, false, true, DIRECT, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, 0u, A, out, gq(per - 2u));
  k3_pp<false, false, true, DIRECT, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, pg, tile_s, lt, 0u, A, out, gq(per - 1u));
  if (DIRECT) { __stcs(out.gbase + (out.g0 ^ pg), double2{pr0, pi0}); __stcs(out.gbase + (out.g1 ^ pg), double2{pr1, pi1}); }
  else { __syncwarp(); xsts<ABL>(pa0, pr0, pi0); xsts<ABL>(pa1, pr1, pi1); }
  tr(8);
}

// One warp's share of a three-product round on the tile at shared address `tile_s`.  A holds variant `cur` on entry.
template <bool DIRECT, class F, class FF>
__device__ __forceinline__ void k3_round_run(uint32_t tile_s, const uint4* lane_tab_r, const uint32_t* btab, uint32_t per,
                                             uint32_t var_hi, const double* __restrict__ mats, uint32_t lane, double (&A)[12], uint32_t cur,
                                             K3Out out, F&& mid, FF&& far_fix, const TileTrace tr = TileTrace{}) {
  const uint4 lt = lane_tab_r[2u * lane];
  out.tile_s = tile_s; out.lz = lt.z; out.lw = lt.w;
  if (DIRECT) {   // lane parts of the global offsets: second table entry of the lane (written by the prologue for the last round)
    const uint4 lg = lane_tab_r[2u * lane + 1u];
    out.g0 = (uint64_t)lg.x | ((uint64_t)lg.y << 32); out.g1 = (uint64_t)lg.z | ((uint64_t)lg.w << 32);
  }
  if ((btab[0] >> 20) == (btab[per - 1u] >> 20)) {
    // local condition bits are the top bits of the batch index: equal at both ends => one variant for the whole share
#ifndef QCB_K3_ROTATE
    if (per >= 4u && !(per & 1u)) {
#ifdef QCB_TILE_ABLATE
      switch (tile_dbg() & 3) {
        case 1: k3_batches_pp<DIRECT, 1>(tile_s, lt, btab, per, A, out, mid); return;
        case 2: k3_batches_pp<DIRECT, 2>(tile_s, lt, btab, per, A, out, mid); return;
        case 3: k3_batches_pp<DIRECT, 3>(tile_s, lt, btab, per, A, out, mid); return;
        default: break;
      }
#endif
      k3_batches_pp<DIRECT>(tile_s, lt, btab, per, A, out, mid, tr);
      return;
    }
#endif
    mid();
    k3_batches<DIRECT>(tile_s, lt, btab, per, A, out);
    return;
  }
  mid();
  // the variant changes inside the share: runs of equal variants, each through the pipelined loop
  uint32_t b = 0;
  while (b < per) {
    const uint32_t v = var_hi | (btab[b] >> 20);
    uint32_t e = b + 1u;
    while (e < per && (btab[e] >> 20) == (btab[b] >> 20)) ++e;
    if (v != cur) { k3_load_A(A, mats, v); far_fix(A); cur = v; }
    K3Out o2 = out;
    o2.gtab = out.gtab + b;
#ifndef QCB_K3_ROTATE
    if (e - b >= 4u && !((e - b) & 1u)) k3_batches_pp<DIRECT>(tile_s, lt, btab + b, e - b, A, o2, [] {});
    else
#endif
    k3_batches<DIRECT>(tile_s, lt, btab + b, e - b, A, o2);
    b = e;
  }
}

// ------------------------------------------------------------------ paired rounds (round kind 3, tile_core.h "paired rounds")
// Two dense 8x8 complex blocks per pass: the first block's D registers ARE the second block's B registers (the fragment
// layouts of mma.m8n8k4 transpose lane-group bits and pattern bits for free), so a batch costs the same two 16-byte loads and
// two 16-byte stores as a single round but carries twelve DMMA.8x8x4 - half the shared-memory traffic per tensor instruction.
// A[0..5] = P0 P1 N0 N1 R0 R1 of the first block, A[6..11] of the second.
__device__ __forceinline__ void k3x_load_A(double (&A)[12], const double* __restrict__ mats, uint32_t var) {
#pragma unroll
  for (int i = 0; i < 12; ++i) A[i] = __ldg(mats + (size_t)var * K3X_FRAG_DOUBLES + i * 32);
}
struct K3XSet {
  double r0, i0, r1, i1;             // raw loads = operands of the first block
  double s0, s1, d0, d1;             // Br + Bi, Bi - Br of the block in progress
  double K0, K1;                     // K = P Br of the block in progress
  double x0, x1, y0, y1;             // results of the first block (Re, Im of columns 0, 1) = operands of the second
  uint32_t X;                        // swizzled byte offset of the batch the set currently belongs to
};
// first block of the batch in `n`, on its own (pipeline prologue)
template <int ABL = 0>
__device__ __forceinline__ void k3x_first_block(K3XSet& n, const double (&A)[12]) {
  n.s0 = n.r0 + n.i0; n.d0 = n.i0 - n.r0; n.s1 = n.r1 + n.i1; n.d1 = n.i1 - n.r1;
  xmma<ABL>(n.K0, n.K1, A[0], n.r0, 0.0, 0.0);
  xmma<ABL>(n.K0, n.K1, A[1], n.r1, n.K0, n.K1);
  xmma<ABL>(n.x0, n.x1, A[2], n.s0, n.K0, n.K1);
  xmma<ABL>(n.y0, n.y1, A[4], n.d0, n.K0, n.K1);
  xmma<ABL>(n.x0, n.x1, A[3], n.s1, n.x0, n.x1);
  xmma<ABL>(n.y0, n.y1, A[5], n.d1, n.y0, n.y1);
}
// One batch of the steady state: the SECOND block of batch i (set c: x / y ready) interleaved with the FIRST block of batch
// i+1 (set n: raw loads landed), so that dependent tensor instructions are at least two issue slots apart; the loads of batch
// i+2 go into c's raw registers (dead since the previous call), the results of batch i-1 are stored at the top.
template <bool HAS1, bool HAS2, bool HASP, int ABL = 0>
__device__ __forceinline__ void k3x_pp(K3XSet& c, K3XSet& n, double& pr0, double& pr1, double& pi0, double& pi1, uint32_t& pa0, uint32_t& pa1,
                                       uint32_t tile_s, const uint4& lt, uint32_t xq, const double (&A)[12]) {
  if (HASP) {
    __syncwarp();                      // in-place update inside a warp: see k3_pp
    xsts<ABL>(pa0, pr0, pi0); xsts<ABL>(pa1, pr1, pi1);
  }
  pa0 = tile_s + (lt.z ^ c.X); pa1 = tile_s + (lt.w ^ c.X);
  if (HAS2) {
    c.X = xq & DMMA_BATCH_OFF_MASK;
    xlds<ABL>(tile_s + (lt.x ^ c.X), c.r0, c.i0);
    xlds<ABL>(tile_s + (lt.y ^ c.X), c.r1, c.i1);
  }
  xmma<ABL>(c.K0, c.K1, A[6], c.x0, 0.0, 0.0);
  if (HAS1) xmma<ABL>(n.K0, n.K1, A[0], n.r0, 0.0, 0.0);
  xmma<ABL>(c.K0, c.K1, A[7], c.x1, c.K0, c.K1);
  if (HAS1) xmma<ABL>(n.K0, n.K1, A[1], n.r1, n.K0, n.K1);
  c.s0 = c.x0 + c.y0; c.d0 = c.y0 - c.x0;
  if (HAS1) { n.s0 = n.r0 + n.i0; n.d0 = n.i0 - n.r0; }
  xmma<ABL>(pr0, pr1, A[8], c.s0, c.K0, c.K1);
  if (HAS1) xmma<ABL>(n.x0, n.x1, A[2], n.s0, n.K0, n.K1);
  xmma<ABL>(pi0, pi1, A[10], c.d0, c.K0, c.K1);
  if (HAS1) xmma<ABL>(n.y0, n.y1, A[4], n.d0, n.K0, n.K1);
  c.s1 = c.x1 + c.y1; c.d1 = c.y1 - c.x1;
  if (HAS1) { n.s1 = n.r1 + n.i1; n.d1 = n.i1 - n.r1; }
  xmma<ABL>(pr0, pr1, A[9], c.s1, pr0, pr1);
  if (HAS1) xmma<ABL>(n.x0, n.x1, A[3], n.s1, n.x0, n.x1);
  xmma<ABL>(pi0, pi1, A[11], c.d1, pi0, pi1);
  if (HAS1) xmma<ABL>(n.y0, n.y1, A[5], n.d1, n.y0, n.y1);
}
// per even and >= 4, one matrix variant.  (A lockstep variant - two batches through the first block together, then through the
// second - measured the same: 5485 vs 5511 gates/s, profiles/r2n_ab.log; on registers alone it loses 9 % to the drain between
// the blocks, scripts/dmma_mix.cu.)
template <int ABL = 0, class F>
__device__ __forceinline__ void k3x_batches_pp(uint32_t tile_s, const uint4 lt, const uint32_t* btab, uint32_t per, const double (&A)[12], F&& mid,
                                               const TileTrace tr = TileTrace{}) {
  K3XSet a, b;
  double pr0 = 0, pr1 = 0, pi0 = 0, pi1 = 0;
  uint32_t pa0 = 0, pa1 = 0;
  a.X = btab[0] & DMMA_BATCH_OFF_MASK;
  b.X = btab[1] & DMMA_BATCH_OFF_MASK;
  xlds<ABL>(tile_s + (lt.x ^ a.X), a.r0, a.i0);
  xlds<ABL>(tile_s + (lt.y ^ a.X), a.r1, a.i1);
  xlds<ABL>(tile_s + (lt.x ^ b.X), b.r0, b.i0);
  xlds<ABL>(tile_s + (lt.y ^ b.X), b.r1, b.i1);
  uint32_t xq = btab[2];
  tr(4);
  k3x_first_block<ABL>(a, A);
  k3x_pp<true, true, false, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, xq, A);
  xq = btab[3];
  k3x_pp<true, true, true, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, xq, A);
  tr(5);
  mid();
  tr(6);
#pragma unroll 1
  for (uint32_t i = 2; i + 2u < per; i += 2u) {
    xq = btab[i + 2u];
    k3x_pp<true, true, true, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, xq, A);
    xq = btab[i + 3u];
    k3x_pp<true, true, true, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, xq, A);
  }
  tr(7);
  k3x_pp<true, false, true, ABL>(a, b, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, 0u, A);
  k3x_pp<false, false, true, ABL>(b, a, pr0, pr1, pi0, pi1, pa0, pa1, tile_s, lt, 0u, A);
  __syncwarp();
  xsts<ABL>(pa0, pr0, pi0); xsts<ABL>(pa1, pr1, pi1);
  tr(8);
}
// any number of batches, one after the other (short shares, variant changes inside a share)
__device__ __forceinline__ void k3x_batches_simple(uint32_t tile_s, const uint4 lt, const uint32_t* btab, uint32_t per, const double (&A)[12]) {
#pragma unroll 1
  for (uint32_t i = 0; i < per; ++i) {
    K3XSet t;
    t.X = btab[i] & DMMA_BATCH_OFF_MASK;
    lds_c128(tile_s + (lt.x ^ t.X), t.r0, t.i0);
    lds_c128(tile_s + (lt.y ^ t.X), t.r1, t.i1);
    k3x_first_block<0>(t, A);
    double re0, re1, im0, im1;
    t.s0 = t.x0 + t.y0; t.d0 = t.y0 - t.x0; t.s1 = t.x1 + t.y1; t.d1 = t.y1 - t.x1;
    dmma_884_c(t.K0, t.K1, A[6], t.x0, 0.0, 0.0);
    dmma_884_c(t.K0, t.K1, A[7], t.x1, t.K0, t.K1);
    dmma_884_c(re0, re1, A[8], t.s0, t.K0, t.K1);
    dmma_884_c(im0, im1, A[10], t.d0, t.K0, t.K1);
    dmma_884_c(re0, re1, A[9], t.s1, re0, re1);
    dmma_884_c(im0, im1, A[11], t.d1, im0, im1);
    __syncwarp();
    sts_c128(tile_s + (lt.z ^ t.X), re0, im0);
    sts_c128(tile_s + (lt.w ^ t.X), re1, im1);
  }
}
// One warp's share of a paired round on the tile at shared address `tile_s`.  A holds variant `cur` on entry.
template <class F, class FF>
__device__ __forceinline__ void k3x_round_run(uint32_t tile_s, const uint4* lane_tab_r, const uint32_t* btab, uint32_t per,
                                              uint32_t var_hi, const double* __restrict__ mats, uint32_t lane, double (&A)[12], uint32_t cur,
                                              F&& mid, FF&& far_fix, const TileTrace tr = TileTrace{}) {
  const uint4 lt = lane_tab_r[2u * lane];
  if ((btab[0] >> 20) == (btab[per - 1u] >> 20)) {
    if (per >= 4u && !(per & 1u)) {
#ifdef QCB_TILE_ABLATE
      switch (tile_dbg() & 3) {
        case 1: k3x_batches_pp<1>(tile_s, lt, btab, per, A, mid); return;
        case 2: k3x_batches_pp<2>(tile_s, lt, btab, per, A, mid); return;
        case 3: k3x_batches_pp<3>(tile_s, lt, btab, per, A, mid); return;
        default: break;
      }
#endif
      k3x_batches_pp<0>(tile_s, lt, btab, per, A, mid, tr);
    } else { mid(); k3x_batches_simple(tile_s, lt, btab, per, A); }
    return;
  }
  mid();
  uint32_t b = 0;
  while (b < per) {
    const uint32_t v = var_hi | (btab[b] >> 20);
    uint32_t e = b + 1u;
    while (e < per && (btab[e] >> 20) == (btab[b] >> 20)) ++e;
    if (v != cur) { k3x_load_A(A, mats, v); far_fix(A); cur = v; }
    if (e - b >= 4u && !((e - b) & 1u)) k3x_batches_pp<0>(tile_s, lt, btab + b, e - b, A, [] {});
    else k3x_batches_simple(tile_s, lt, btab + b, e - b, A);
    b = e;
  }
}

// Pacing of the mover (QCB_MOVER_PAUSE_NS, experiment knob): nanoseconds slept between groups of 8 element copies so that a
// burst of mover LDGSTS / LDS.128 does not monopolise the LSU pipe the consumers' fragment loads depend on.  0 = off.
__device__ unsigned g_mover_pause_ns;

// ------------------------------------------------------------------ LSU mover, specialised
// Tiles of 128 * KE amplitudes with 256-byte runs and the default layout (the common case): mover thread mt moves
// elements mt + 128 k (k < KE); their global run offsets stay in registers for the whole sweep and the swizzled
// shared-memory offsets are XORs with compile-time constants (swz is linear), so one element costs an address add and
// one LDGSTS / LDS + STG.
template <int KE>
__device__ __forceinline__ void mover_fast(double2* __restrict__ state, const uint64_t* sprog, const StageCtx& sc, const uint64_t* hoff,
                                           uint32_t smem_s, uint32_t tile_bytes, uint32_t mt, uint32_t T, uint32_t nbuf,
                                           uint64_t* full, uint64_t* done, bool direct) {
  const uint32_t low = mt & 15u;
  const uint32_t s_mt = swz(mt, 0u) << 4;
  const unsigned pause = g_mover_pause_ns;
  DBG_DECL;
  uint32_t roff[KE];                                           // run offsets in units of 256 bytes
#pragma unroll
  for (uint32_t k = 0; k < KE; ++k) roff[k] = (uint32_t)(hoff[(mt >> 4) + 8u * k] >> 4);
  PF_DECL;
  if (direct) {
    // the consumers write the results to global memory themselves (last round): this warpgroup only loads, as soon as the
    // buffer's previous tile has been released
    for (uint32_t j = 0; j < T; ++j) {
      if (j >= nbuf) { const uint32_t s = j - nbuf; mbar_wait<true>(done + (s % nbuf), (s / nbuf) & 1u); PF_ADD(PF_M_WAIT_DONE); }
      const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
      const char* gb = reinterpret_cast<const char*>(state + tile_base(sprog, sc, t) + low);
      const uint32_t bs = smem_s + (j % nbuf) * tile_bytes;
#pragma unroll
      for (uint32_t k = 0; k < KE; ++k) cp_async16_s(bs + (s_mt ^ (swz(128u * k, 0u) << 4)), gb + ((uint64_t)roff[k] << 8));
      cp_async_mbar_arrive(full + (j % nbuf));
      PF_ADD(PF_M_LOAD);
    }
    PF_TOTAL(PF_M_TOTAL);
    return;
  }
  for (uint32_t j = 0; j < T + nbuf - 1u; ++j) {
    if (j < T) {
      const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
      const char* gb = reinterpret_cast<const char*>(state + tile_base(sprog, sc, t) + low);
      const uint32_t bs = smem_s + (j % nbuf) * tile_bytes;
      if (!DBG_ON(4)) {
#pragma unroll
        for (uint32_t k = 0; k < KE; ++k) {
          cp_async16_s(bs + (s_mt ^ (swz(128u * k, 0u) << 4)), gb + ((uint64_t)roff[k] << 8));
          if (pause && (k & 7u) == 7u) __nanosleep(pause);
        }
      }
      cp_async_mbar_arrive(full + (j % nbuf));
      TRACE_MAKE(8u + (mt >> 5), j, 0u)(20);
      PF_ADD(PF_M_LOAD);
    }
    if (j + 1u >= nbuf) {
      const uint32_t s = j + 1u - nbuf;                        // tile to write back (s < T by the loop bound)
      mbar_wait<true>(done + (s % nbuf), (s / nbuf) & 1u);
      TRACE_MAKE(8u + (mt >> 5), s, 0u)(21);
      PF_ADD(PF_M_WAIT_DONE);
      const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x);
      char* gb = reinterpret_cast<char*>(state + tile_base(sprog, sc, t) + low);
      const uint32_t bs = smem_s + (s % nbuf) * tile_bytes;
#pragma unroll
      for (uint32_t k0 = 0; k0 < (DBG_ON(4) ? 0u : (uint32_t)KE); k0 += 8) {
        double2 v[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) v[u] = lds_f64x2(bs + (s_mt ^ (swz(128u * (k0 + u), 0u) << 4)));
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) __stcs(reinterpret_cast<double2*>(gb + ((uint64_t)roff[k0 + u] << 8)), v[u]);
        if (pause) __nanosleep(pause);
      }
      TRACE_MAKE(8u + (mt >> 5), s, 0u)(22);
      PF_ADD(PF_M_STORE);
    }
  }
  PF_TOTAL(PF_M_TOTAL);
}

// ------------------------------------------------------------------ the fused gate executor
// One persistent CTA per SM, warp-specialised over a ring of `nbuf` tile buffers in shared memory:
//   mover            streams tile j into buffer j % nbuf (completion signalled on full[]), then writes tile j-(nbuf-1)
//                    back to HBM once its consumer group has signalled done[].  TMA mode (the normal case): ONE warp
//                    issues a cp.async.bulk.tensor copy per contiguous run of the tile, the TMA unit swizzles on the
//                    fly and the LSU pipe stays free for the consumers.  Fallback (tiles smaller than a 128-byte row
//                    or no tensor map): four warps move 16 bytes per thread with cp.async / st.global;
//   consumer groups  NG groups of WPG warps; group g runs the stage's rounds on tiles j = g, g+NG, ... in shared
//                    memory (tensor-core rounds or interpreter rounds), rounds separated by the group's own named
//                    barrier.  Two groups on different tiles keep the fp64 tensor pipe fed while the other group
//                    sits at a round barrier or waits for operands.
// HBM traffic of the tiles ahead (loads) and behind (stores) overlaps the arithmetic.  In the fallback each mover thread
// loads and stores the same elements of a buffer, so buffer re-use needs no further barrier than its own program order;
// in TMA mode the mover waits for its bulk stores to have read the buffer (wait_group.read) before reloading it.
constexpr int MOVER_WARPS = 4;
constexpr int MOVER_THREADS = MOVER_WARPS * 32;

// MMA_ONLY = true: every round of the stage is a tensor-core round (the common case); the op interpreter is
// compiled out, which keeps the hot loop free of register spills.
// FORM: kind of the stage's tensor-core rounds (one per plan): 2 = three-product form (default), 1 = 16x16 real block.
template <int NG, int WPG, bool MMA_ONLY, int FORM>
__global__ void __launch_bounds__((NG * WPG + MOVER_WARPS) * 32, 1)
k_tile_stage(double2* __restrict__ state, const uint64_t* __restrict__ stage_g, uint32_t stage_words,
             const double* __restrict__ dev_vals, uint64_t n_active, uint32_t nbuf, const __grid_constant__ CUtensorMap tmap,
             uint32_t use_tma) {
  // stage_words = descriptor part of the stage program (stage + round descriptors + interpreter op slots);
  // tensor-core matrices follow it in global memory and are read through the read-only path
  extern __shared__ unsigned char smem_dyn[];
  // the hardware swizzle is a function of the shared-memory address: tile buffers start on a 1024-byte boundary
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr uint32_t NCW = NG * WPG, NTHREADS = (NCW + MOVER_WARPS) * 32, NCT = NCW * 32, GT = WPG * 32;
  StageCtx sc;
  decode_stage(stage_g, sc);
  const uint32_t m = sc.m, LC = sc.c, tile_n = 1u << m, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint32_t L = use_tma ? sc.c : sc.L;               // run bits: amplitudes moved per copy (TMA) / per hoff entry (LSU)
  const size_t tile_bytes = (size_t)16 << m;
  const uint32_t nbstride = tile_n >= 64u ? (tile_n >> 6) : 1u;
  unsigned char* p = smem_raw + (size_t)nbuf * tile_bytes;
  uint64_t* sprog = reinterpret_cast<uint64_t*>(p);       p += 8 * (size_t)((stage_words + 1u) & ~1u);
  uint64_t* hoff = reinterpret_cast<uint64_t*>(p);        p += ((size_t)(8u << (m - L)) + 15u) & ~(size_t)15u;
  uint4* lane_tab = reinterpret_cast<uint4*>(p);          p += (size_t)1024 * sc.n_rounds;
  uint32_t* batch_tab = reinterpret_cast<uint32_t*>(p);   p += (((size_t)4 * nbstride * sc.n_rounds) + 15u) & ~(size_t)15u;
  uint32_t* run_dst = reinterpret_cast<uint32_t*>(p);     p += ((size_t)(4u << (m - L)) + 15u) & ~(size_t)15u;
  uint2* rtab = reinterpret_cast<uint2*>(p);              p += ((size_t)8 * sc.n_rounds + 15u) & ~(size_t)15u;   // {hi_desc, matrix word offset}
  uint64_t* gtab = reinterpret_cast<uint64_t*>(p);        p += (size_t)8 * nbstride;                             // direct store: global offset of every batch of the last round
  uint64_t* full = reinterpret_cast<uint64_t*>(p);
  uint64_t* done = full + nbuf;

  for (uint32_t i = tid; i < stage_words; i += NTHREADS) sprog[i] = stage_g[i];
  for (uint32_t i = tid; i < (1u << (m - L)); i += NTHREADS) {
    hoff[i] = hi_offset_from(stage_g, sc, i, L);         // global offset (amplitudes) of run i inside a tile
    run_dst[i] = (swz(i << L, LC) >> 3) << 7;            // byte offset of the first 128-byte row of run i in a tile buffer
  }
  if (tid == 0)
    for (uint32_t b = 0; b < nbuf; ++b) { mbar_init(full + b, use_tma ? 1u : MOVER_THREADS); mbar_init(done + b, WPG); }
  __syncthreads();
  for (uint32_t idx = tid; idx < sc.n_rounds * 32u; idx += NTHREADS) {
    const uint32_t r = idx >> 5, kd = round_kind(sprog, r);
    if (kd != (uint32_t)FORM && !(FORM == 2 && kd == 3u)) continue;
    if (FORM == 2) {
      K3Ctx c;
      decode_k3(sprog, r, c);
      uint32_t e[4];
      if (kd == 3u) k3x_lane_entry(c, idx & 31u, e); else k3_lane_entry(c, idx & 31u, e);
      lane_tab[2u * idx] = make_uint4(e[0], e[1], e[2], e[3]);
    } else {
      DmmaCtx c;
      decode_dmma(sprog, r, c);
      uint32_t e[8];
      dmma_lane_entry(c, idx & 31u, e);
      lane_tab[2u * idx] = make_uint4(e[0], e[1], e[2], e[3]);
      lane_tab[2u * idx + 1u] = make_uint4(e[4], e[5], e[6], e[7]);
    }
  }
  for (uint32_t r = tid; r < sc.n_rounds; r += NTHREADS) {
    const uint64_t* w = sprog + T_STAGE_WORDS + (uint64_t)r * T_ROUND_WORDS;
    uint32_t hd = 0;
    for (uint32_t j = 0; j < 4; ++j) {
      const uint32_t pz = (uint32_t)w[30 + j];
      hd |= ((j < (uint32_t)w[29] && pz >= m) ? (pz - m) : 63u) << (6u * j);
    }
    rtab[r] = make_uint2(hd, (uint32_t)w[2]);
  }
  for (uint32_t idx = tid; idx < sc.n_rounds * nbstride; idx += NTHREADS) {
    const uint32_t r = idx / nbstride, b = idx - r * nbstride, kd = round_kind(sprog, r);
    if (kd != (uint32_t)FORM && !(FORM == 2 && kd == 3u)) continue;
    DmmaCtx c;                                            // the batch geometry words are common to both forms
    decode_dmma(sprog, r, c);
    if (b < (1u << (c.n_grp - 3u))) batch_tab[idx] = dmma_batch_entry(c, b, m);
  }
  // direct store (FORM 2 only): the last round writes to global memory; its batches' and lanes' global offsets
  const bool direct = FORM == 2 && (stage_g[41] & T_FLAG_DIRECT_STORE) != 0 && !use_tma && sc.n_rounds > 0;
  if (FORM == 2 && direct) {
    const uint32_t rl = sc.n_rounds - 1u;
    K3Ctx c;
    decode_k3(sprog, rl, c);
    auto goff = [&](uint32_t idx) -> uint64_t { return hoff[idx >> L] + (uint64_t)(idx & ((1u << L) - 1u)); };
    for (uint32_t b = tid; b < (1u << (c.n_grp - 3u)); b += NTHREADS) gtab[b] = goff(k3_batch_base(c, b));
    if (tid < 32u) {
      const uint64_t a0 = goff(k3_lane_store_index(c, tid, 0u)), a1 = goff(k3_lane_store_index(c, tid, 1u));
      lane_tab[2u * (rl * 32u + tid) + 1u] = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32));
    }
  }
  __syncthreads();

  const uint32_t T = (uint32_t)((n_active - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA
  const uint32_t lowmask = (1u << L) - 1u;
  // Register re-allocation between the warpgroups (setmaxnreg): the kernel is compiled for 168 registers per thread (384
  // threads, one CTA per SM); the mover warpgroup hands registers to the two consumer warpgroups (112 / 192 per thread).
#ifndef QCB_NO_REALLOC
  constexpr bool REALLOC = (NCW == 8 && MOVER_WARPS == 4);      // measured: +6 % (5166 -> 5473 gates/s, profiles/r2a_ab.log)
#else
  constexpr bool REALLOC = false;
#endif

  if (warp >= NCW) {
  if constexpr (REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 112;\n");
  if (use_tma) {
    // ---------------- mover (TMA): one warp, one bulk tensor copy per run
    if (warp > NCW) return;
    const uint32_t nruns = 1u << (m - L), smem_s = smem_u32(smem_raw);
    PF_DECL;
    for (uint32_t j = 0; j < T + nbuf - 1u; ++j) {
      if (j < T) {
        const uint32_t b = j % nbuf;
        const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
        const uint64_t gb = tile_base(sprog, sc, t);
        if (lane == 0) mbar_arrive_expect_tx(full + b, (uint32_t)tile_bytes);
        __syncwarp();
        for (uint32_t h = lane; h < nruns; h += 32u)
          tma_load_rows(smem_s + b * (uint32_t)tile_bytes + run_dst[h], &tmap, (int32_t)((gb + hoff[h]) >> 3), full + b);
        PF_ADD(PF_M_LOAD);
      }
      if (j + 1u >= nbuf) {
        const uint32_t s = j + 1u - nbuf, b = s % nbuf;        // tile to write back (s < T by the loop bound)
        mbar_wait<true>(done + b, (s / nbuf) & 1u);
        PF_ADD(PF_M_WAIT_DONE);
        const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x);
        const uint64_t gb = tile_base(sprog, sc, t);
        for (uint32_t h = lane; h < nruns; h += 32u)
          tma_store_rows(&tmap, (int32_t)((gb + hoff[h]) >> 3), smem_s + b * (uint32_t)tile_bytes + run_dst[h]);
        tma_store_commit();
        tma_store_wait_read();                                 // the buffer has been read: it may be reloaded
        __syncwarp();
        PF_ADD(PF_M_STORE);
      }
    }
    PF_TOTAL(PF_M_TOTAL);
  } else if (L == 4u && LC == 0u && (m == 12u || m == 11u)) {
    // ---------------- mover warps, 64 KB / 32 KB tiles with 256-byte runs (the common case)
    if (m == 12u) mover_fast<32>(state, sprog, sc, hoff, smem_u32(smem_raw), (uint32_t)tile_bytes, tid - NCT, T, nbuf, full, done, direct);
    else mover_fast<16>(state, sprog, sc, hoff, smem_u32(smem_raw), (uint32_t)tile_bytes, tid - NCT, T, nbuf, full, done, false);
  } else {
    // ---------------- mover warps (fallback: 16 bytes per thread through the LSU)
    const uint32_t mt = tid - NCT;
    PF_DECL;
    for (uint32_t j = 0; j < T + nbuf - 1u; ++j) {
      if (j < T) {
        const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
        const double2* gbase = state + tile_base(sprog, sc, t);
        double2* buf = reinterpret_cast<double2*>(smem_raw + (size_t)(j % nbuf) * tile_bytes);
        for (uint32_t i = mt; i < tile_n; i += MOVER_THREADS)
          cp_async16(&buf[swz(i, LC)], gbase + hoff[i >> L] + (i & lowmask));
        cp_async_mbar_arrive(full + (j % nbuf));
        PF_ADD(PF_M_LOAD);
      }
      if (j + 1u >= nbuf) {
        const uint32_t s = j + 1u - nbuf;                      // tile to write back (s < T by the loop bound)
        mbar_wait<true>(done + (s % nbuf), (s / nbuf) & 1u);
        PF_ADD(PF_M_WAIT_DONE);
        const uint64_t t = active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x);
        double2* gbase = state + tile_base(sprog, sc, t);
        const double2* buf = reinterpret_cast<const double2*>(smem_raw + (size_t)(s % nbuf) * tile_bytes);
        for (uint32_t i0 = mt; i0 < tile_n; i0 += MOVER_THREADS * 8u) {
          double2 v[8];
#pragma unroll
          for (uint32_t u = 0; u < 8; ++u) {
            const uint32_t i = i0 + u * MOVER_THREADS;
            if (i < tile_n) v[u] = buf[swz(i, LC)];
          }
#pragma unroll
          for (uint32_t u = 0; u < 8; ++u) {
            const uint32_t i = i0 + u * MOVER_THREADS;
            if (i < tile_n) __stcs(gbase + hoff[i >> L] + (i & lowmask), v[u]);
          }
        }
        PF_ADD(PF_M_STORE);
      }
    }
    PF_TOTAL(PF_M_TOTAL);
  }
  } else {
    // ---------------- consumer groups
    if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;\n");
    const uint32_t grp = warp / WPG, gwarp = warp - grp * WPG, gtid = tid - grp * GT;
    const uint32_t smem_s = smem_u32(smem_raw);
    // every tensor-core round has m - 3 group bits: the batch geometry of this warp is a kernel constant
    const uint32_t nbatch = tile_n >= 64u ? (tile_n >> 6) : 1u;
    const uint32_t per = nbatch >= (uint32_t)WPG ? nbatch / WPG : 1u, b0 = gwarp * per;
    const bool active = b0 < nbatch;
    constexpr int NA = FORM == 2 ? 12 : 8;                // A-fragment registers of one matrix variant (12: a paired round)
    double A[NA];
    uint32_t cur = 0xffffffffu, var_hi = 0;
    DBG_DECL;
    // a tensor-core round of this kernel's form (form 2: single three-product rounds and paired rounds)
    auto is_mma = [&](uint32_t r) {
      const uint32_t kd = round_kind(sprog, r);
      return kd == (uint32_t)FORM || (FORM == 2 && kd == 3u);
    };
    // operands of (tile j, round r): variant bits from the tile id, first variant of this warp, its A fragments
    auto prefetch = [&](uint32_t j, uint32_t r) {
      if (!active) return;
      const uint64_t ext_hi = sc.ext_hi_base | active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
      const uint2 rt = rtab[r];
      var_hi = dmma_var_hi(rt.x, ext_hi);
      cur = var_hi | (batch_tab[r * nbstride + b0] >> 20);
      if constexpr (FORM == 2) {
        if (round_kind(sprog, r) == 3u) k3x_load_A(A, reinterpret_cast<const double*>(stage_g + rt.y) + lane, cur);
        else k3_load_A(A, reinterpret_cast<const double*>(stage_g + rt.y) + lane, cur);
      } else dmma_load_A(A, reinterpret_cast<const double*>(stage_g + rt.y) + lane, cur);
    };
    if (grp < T && is_mma(0)) prefetch(grp, 0);
    PF_DECL;
    for (uint32_t j = grp; j < T; j += NG) {
      const uint32_t b = j % nbuf;
      TRACE_MAKE(warp, j, 0u)(0);
      mbar_wait(full + b, (j / nbuf) & 1u);
      TRACE_MAKE(warp, j, 0u)(1);
      PF_ADD(PF_C_WAIT_FULL);
      for (uint32_t r = 0; r < sc.n_rounds; ++r) {
        const TileTrace tr = TRACE_MAKE(warp, j, r);
        tr(9);
        if (r && !DBG_ON(8)) group_bar_sync<GT>(grp);
        tr(2);
        PF_ADD(PF_C_BARRIER);
        uint32_t nj = j, nr = r + 1u;
        if (nr == sc.n_rounds) { nr = 0; nj = j + NG; }
        const bool next_mma = nj < T && (MMA_ONLY || is_mma(nr));
        if (MMA_ONLY || is_mma(r)) {
          double Ac[NA];
#pragma unroll
          for (int i = 0; i < NA; ++i) Ac[i] = A[i];
          const uint32_t curc = cur, var_hic = var_hi;
          const double* mats = reinterpret_cast<const double*>(stage_g + rtab[r].y) + lane;
          // far phases (tile_core.h): diagonal gates with an operand outside the tile scale the rows (gates after the block) or
          // the columns (gates before it) of the block by constants of the tile - applied to a freshly loaded set of fragments
          auto far_fix = [&](double (&X)[12]) {
            if constexpr (FORM == 2) {
              const uint64_t* w = sprog + T_STAGE_WORDS + (uint64_t)r * T_ROUND_WORDS;
              const uint32_t fn = (uint32_t)w[39];
              if (fn == 0u) return;
              const uint64_t ext_hi = sc.ext_hi_base | active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
              const uint32_t mm1[3] = {(uint32_t)w[35] & 15u, (uint32_t)(w[35] >> 4) & 15u, (uint32_t)(w[35] >> 8) & 15u};
              const uint32_t km1[3] = {(uint32_t)w[34] & 15u, (uint32_t)(w[34] >> 4) & 15u, (uint32_t)(w[34] >> 8) & 15u};
              const uint32_t mm2[3] = {(uint32_t)w[36] & 15u, (uint32_t)(w[36] >> 4) & 15u, (uint32_t)(w[36] >> 8) & 15u};
              double fs[4], dr, di;
#pragma unroll
              for (int blk = 0; blk < 2; ++blk) {
                const uint32_t n_post = (fn >> (8 * blk)) & 0xffu, n_pre = (fn >> (16 + 8 * blk)) & 0xffu;
                const uint64_t offs = w[37 + blk];
                if (n_post) {
                  far_sums_warp(stage_g + (uint32_t)offs, n_post, ext_hi, lane, fs);
                  sincos(far_angle(fs, k3_pattern_index(lane >> 2, blk ? mm2 : mm1)), &di, &dr);
                  far_scale(X[6 * blk + 0], X[6 * blk + 2], X[6 * blk + 4], dr, di);
                  far_scale(X[6 * blk + 1], X[6 * blk + 3], X[6 * blk + 5], dr, di);
                }
                if (n_pre) {
                  far_sums_warp(stage_g + (uint32_t)(offs >> 32), n_pre, ext_hi, lane, fs);
#pragma unroll
                  for (uint32_t s2 = 0; s2 < 2; ++s2) {
                    const uint32_t kcol = (lane & 3u) + 4u * s2;
                    sincos(far_angle(fs, blk ? k3x_hw_k_to_group(kcol) : k3_pattern_index(kcol, km1)), &di, &dr);
                    far_scale(X[6 * blk + s2], X[6 * blk + 2 + s2], X[6 * blk + 4 + s2], dr, di);
                  }
                }
              }
            }
          };
          if constexpr (FORM == 2) { if (active) far_fix(Ac); }
          const bool do_pf = next_mma && !(DBG_ON(16) && cur != 0xffffffffu);
#ifdef QCB_PREFETCH_EARLY
          if (do_pf) prefetch(nj, nr);
          auto mid = [] {};
#else
          auto mid = [&] { if (do_pf) prefetch(nj, nr); };
          if ((FORM != 2 || !active) && do_pf) prefetch(nj, nr);      // the 16x16 form keeps the prefetch between the passes
#endif
          PF_ADD(PF_C_SETUP);
          tr(3);
          if (active) {
            if constexpr (FORM == 2) {
              if (round_kind(sprog, r) == 3u) {
                k3x_round_run(smem_s + b * (uint32_t)tile_bytes, lane_tab + (size_t)r * 64u, batch_tab + r * nbstride + b0, per, var_hic,
                              mats, lane, Ac, curc, mid, far_fix, tr);
              } else {
              K3Out out;
              out.gtab = gtab + b0; out.gbase = nullptr; out.g0 = out.g1 = 0;
              if (direct && r + 1u == sc.n_rounds) {
                out.gbase = state + tile_base(sprog, sc, active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x));
                k3_round_run<true>(smem_s + b * (uint32_t)tile_bytes, lane_tab + (size_t)r * 64u, batch_tab + r * nbstride + b0, per, var_hic,
                                   mats, lane, Ac, curc, out, mid, far_fix);
              } else {
                k3_round_run<false>(smem_s + b * (uint32_t)tile_bytes, lane_tab + (size_t)r * 64u, batch_tab + r * nbstride + b0, per, var_hic,
                                    mats, lane, Ac, curc, out, mid, far_fix, tr);
              }
              }
            } else
              dmma_round_run(smem_s + b * (uint32_t)tile_bytes, lane_tab + (size_t)r * 64u, batch_tab + r * nbstride + b0, per, var_hic,
                             mats, lane, Ac, curc, dbg_bits);
          }
          PF_ADD(PF_C_ROUND);
        } else if (!MMA_ONLY) {
          RoundCtx rc;
          decode_round(sprog, r, rc);
          const uint64_t ext_hi = sc.ext_hi_base | active_to_tile(sc, (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x);
          double2* t2 = reinterpret_cast<double2*>(smem_raw + (size_t)b * tile_bytes);
          switch (rc.r) {
            case 0: run_round_thread<0>(t2, rc, m, LC, ext_hi, gtid, GT, dev_vals); break;
            case 1: run_round_thread<1>(t2, rc, m, LC, ext_hi, gtid, GT, dev_vals); break;
            case 2: run_round_thread<2>(t2, rc, m, LC, ext_hi, gtid, GT, dev_vals); break;
            default: run_round_thread<3>(t2, rc, m, LC, ext_hi, gtid, GT, dev_vals); break;
          }
          if (next_mma) prefetch(nj, nr);
        }
      }
      if (use_tma) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(done + b);
      TRACE_MAKE(warp, j, 15u)(10);
    }
    PF_TOTAL(PF_C_TOTAL);
  }
}

template <int NG, int WPG>
static cudaError_t launch_tile_stage_t(bool mma_only, int form, unsigned grid, size_t smem, size_t limit, cudaStream_t stream, double2* state,
                                       const uint64_t* stage_dev, uint32_t stage_words, const double* dev_vals, uint64_t n_active,
                                       uint32_t nbuf, const CUtensorMap& tmap, uint32_t use_tma) {
  // function attributes are per device: handles on different GPUs may live in one process
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t dev_bit = 1ULL << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & dev_bit)) {
    cudaError_t e = cudaFuncSetAttribute(k_tile_stage<NG, WPG, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile_stage<NG, WPG, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile_stage<NG, WPG, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile_stage<NG, WPG, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e != cudaSuccess) return e;
    configured.fetch_or(dev_bit, std::memory_order_release);
  }
  const unsigned threads = (NG * WPG + MOVER_WARPS) * 32;
#define QCB_TS_ARGS <<<grid, threads, smem, stream>>>(state, stage_dev, stage_words, dev_vals, n_active, nbuf, tmap, use_tma)
  if (form == 1) { if (mma_only) k_tile_stage<NG, WPG, true, 1> QCB_TS_ARGS; else k_tile_stage<NG, WPG, false, 1> QCB_TS_ARGS; }
  else { if (mma_only) k_tile_stage<NG, WPG, true, 2> QCB_TS_ARGS; else k_tile_stage<NG, WPG, false, 2> QCB_TS_ARGS; }
#undef QCB_TS_ARGS
  return cudaGetLastError();
}

cudaError_t launch_tile_stage(double2* state, const uint64_t* stage_dev, const uint64_t* stage_host, uint32_t stage_words,
                              const double* dev_vals, int num_sms, cudaStream_t stream, uint64_t* out_active, const TileMaps* maps) {
  StageCtx sc;
  decode_stage(stage_host, sc);
  bool mma_only = sc.n_rounds > 0;
  int form = 2;                                          // kind of the tensor-core rounds (the same for every round of a plan)
  for (uint32_t r = 0; r < sc.n_rounds; ++r) {
    const uint32_t kd = round_kind(stage_host, r);
    mma_only = mma_only && kd != 0u;
    if (kd == 1u) form = 1;
  }
  const uint32_t nb = sc.n_local - sc.m;
  const uint64_t tmask = (nb >= 64) ? ~0ULL : ((1ULL << nb) - 1ULL);
  // the part of the skip condition living in the rank bits is decided here, per rank
  if ((sc.ext_hi_base & sc.skip_mask & ~tmask) != (sc.skip_val & ~tmask)) { if (out_active) *out_active = 0; return cudaSuccess; }
  const uint64_t n_active = (1ULL << nb) >> __builtin_popcountll(sc.skip_mask & tmask);
  if (out_active) *out_active = n_active;
  // TMA mode needs a tensor map whose box is one run of 2^c amplitudes (c >= 3: at least one 128-byte row)
  {
    // experiment knobs live in per-device __device__ symbols: set them once per device
    static std::atomic<uint64_t> knobs_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t dev_bit = 1ULL << (dev & 63);
    if (!(knobs_set.load(std::memory_order_acquire) & dev_bit)) {
#ifdef QCB_TILE_PROFILE
      { const char* e = getenv("QCB_TILE_DBG"); int v = e ? atoi(e) : 0; cudaMemcpyToSymbol(g_tile_dbg, &v, sizeof v); }
#endif
      { const char* e = getenv("QCB_MOVER_PAUSE_NS"); unsigned v = e ? (unsigned)atoi(e) : 0u; cudaMemcpyToSymbol(g_mover_pause_ns, &v, sizeof v); }
      knobs_set.fetch_or(dev_bit, std::memory_order_release);
    }
  }
  static const bool no_tma = getenv("QCB_NO_TMA") != nullptr;
  static const CUtensorMap dummy_map = {};
  const uint32_t use_tma = (!no_tma && maps && sc.c >= 3 && sc.c <= 11 && maps->valid[sc.c]) ? 1u : 0u;
  const CUtensorMap* tm = use_tma ? reinterpret_cast<const CUtensorMap*>(maps->map[sc.c]) : &dummy_map;
  const uint32_t rb = use_tma ? sc.c : sc.L;            // run bits (kernel: L)
  const size_t tile_n = (size_t)1 << sc.m, nbstride = tile_n >= 64 ? (tile_n >> 6) : 1;
  const size_t fixed = 8 * (size_t)((stage_words + 1u) & ~1u) + ((((size_t)8 << (sc.m - rb)) + 15) & ~(size_t)15) +
                       (size_t)1024 * sc.n_rounds + ((4 * nbstride * sc.n_rounds + 15) & ~(size_t)15) +
                       ((((size_t)4 << (sc.m - rb)) + 15) & ~(size_t)15) + (((size_t)8 * sc.n_rounds + 15) & ~(size_t)15) + 8 * nbstride +
                       16 * 8 + 1024;   // + run / round tables, direct-store batch offsets, mbarriers (nbuf <= 8), alignment slack
  const size_t limit = 227 * 1024;
  uint64_t grid = (uint64_t)num_sms;
  if (grid > n_active) grid = n_active;
  const uint64_t tiles_per_cta = (n_active + grid - 1) / grid;
  static const int max_buf = [] { const char* e = getenv("QCB_TILE_BUFFERS"); return e ? atoi(e) : 8; }();
  uint32_t nbuf = (uint32_t)(max_buf < 1 ? 1 : (max_buf > 8 ? 8 : max_buf));
  while (nbuf > 1 && (fixed + nbuf * (tile_n * 16) > limit || nbuf > tiles_per_cta + 1)) --nbuf;
  const size_t smem = fixed + nbuf * (tile_n * 16);
  if (smem > limit) return cudaErrorInvalidConfiguration;
  // consumer layout: groups x warps-per-group (QCB_CONSUMERS = "2x4" default, "2x8", "1x8"; 1x16 and 3x4 were measured
  // and dropped: profiles/r1c_sweep_pipelined_groups.log, profiles/r1d_layout_sweep.log)
  static const int layout = [] {
    const char* e = getenv("QCB_CONSUMERS");
    if (!e) return 24;
    return (e[0] - '0') * 10 + atoi(e + 2);
  }();
#define QCB_LAUNCH(NG, WPG) \
  return launch_tile_stage_t<NG, WPG>(mma_only, form, (unsigned)grid, smem, limit, stream, state, stage_dev, stage_words, dev_vals, n_active, nbuf, \
                                      *tm, use_tma)
  switch (layout) {
    case 18: QCB_LAUNCH(1, 8);
    case 28: QCB_LAUNCH(2, 8);
    default: QCB_LAUNCH(2, 4);
  }
#undef QCB_LAUNCH
}

// ------------------------------------------------------------------ tensor maps of the state (host)
cudaError_t build_tile_maps(double2* state, int n_local, TileMaps* out) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is a 128-byte opaque object");
  for (int c = 0; c < 12; ++c) out->valid[c] = false;
  if (n_local < 3) return cudaSuccess;                     // smaller than one 128-byte row: LSU mover only
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return e;
  if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  const cuuint64_t rows = 1ULL << (n_local - 3);
  for (int c = 3; c <= 11 && c <= n_local; ++c) {
    const cuuint64_t gdim[2] = {16, rows};                 // 16 doubles = 8 amplitudes = 128 bytes per row
    const cuuint64_t gstride[1] = {128};                   // bytes between rows
    const cuuint32_t box[2] = {16, 1u << (c - 3)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(reinterpret_cast<CUtensorMap*>(out->map[c]), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, state,
                                                gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    out->valid[c] = (r == CUDA_SUCCESS);
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------ state initialisation
__global__ void k_set_amp(double2* state, uint64_t idx, double re, double im) { state[idx] = double2{re, im}; }

cudaError_t launch_set_amp(double2* state, uint64_t idx, double re, double im, cudaStream_t s) {
  k_set_amp<<<1, 1, 0, s>>>(state, idx, re, im);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ reductions
// mode 0: sum of amplitudes (re, im); mode 1: sum |a|^2 (out[0]).  partials: [grid][2]
__global__ void __launch_bounds__(RED_THREADS)
k_reduce(const double2* __restrict__ state, uint64_t count, int mode, double* __restrict__ partials) {
  __shared__ double sm[2 * 32];
  double v[2] = {0.0, 0.0};
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  // streaming read: four independent 16-byte loads in flight per thread (one load per iteration leaves the memory
  // system latency-bound at ~0.73 of the copy bandwidth, profiles/r1e_launches.csv)
  for (; i + 3 * stride < count; i += 4 * stride) {
    const double2 a0 = __ldcs(state + i), a1 = __ldcs(state + i + stride), a2 = __ldcs(state + i + 2 * stride), a3 = __ldcs(state + i + 3 * stride);
    if (mode == 0) { v[0] += (a0.x + a1.x) + (a2.x + a3.x); v[1] += (a0.y + a1.y) + (a2.y + a3.y); }
    else { v[0] += (a0.x * a0.x + a0.y * a0.y) + (a1.x * a1.x + a1.y * a1.y) + ((a2.x * a2.x + a2.y * a2.y) + (a3.x * a3.x + a3.y * a3.y)); }
  }
  for (; i < count; i += stride) {
    const double2 a = __ldcs(state + i);
    if (mode == 0) { v[0] += a.x; v[1] += a.y; }
    else { v[0] += a.x * a.x + a.y * a.y; }
  }
  block_sum<2>(v, sm);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = v[0]; partials[2 * blockIdx.x + 1] = v[1]; }
}

// Sums `nparts` partial vectors of K doubles in a fixed order; one block.
// post: 0 = plain; 1 = Grover reflection coefficients -> out = {-1, 0, 2*re/N, 2*im/N} (N = param);
//       2 = normalisation coefficients -> out = {1/sqrt(s), 0, 0, 0} when sqrt(s) > tol (param) else {1,0,0,0}
__global__ void __launch_bounds__(RED_THREADS)
k_finalize(const double* __restrict__ partials, uint32_t nparts, uint32_t K, int post, double param, double* __restrict__ out) {
  __shared__ double sm[32];
  for (uint32_t k = 0; k < K; ++k) {
    double v[1] = {0.0};
    for (uint32_t p = threadIdx.x; p < nparts; p += RED_THREADS) v[0] += partials[(uint64_t)p * K + k];
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) {
      if (post == 0) out[k] = v[0];
      else sm[16 + k] = v[0];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (post == 1) { out[0] = -1.0; out[1] = 0.0; out[2] = 2.0 * sm[16] / param; out[3] = 2.0 * sm[17] / param; }
    else if (post == 2) {
      const double nrm = sqrt(sm[16]);
      out[0] = (nrm > 0.0 && nrm > param) ? 1.0 / nrm : 1.0; out[1] = 0.0; out[2] = 0.0; out[3] = 0.0;
    }
  }
}

cudaError_t launch_reduce(const double2* state, uint64_t count, int mode, double* partials, int grid, cudaStream_t s) {
  k_reduce<<<grid, RED_THREADS, 0, s>>>(state, count, mode, partials);
  return cudaGetLastError();
}
cudaError_t launch_finalize(const double* partials, uint32_t nparts, uint32_t K, int post, double param, double* out, cudaStream_t s) {
  k_finalize<<<1, RED_THREADS, 0, s>>>(partials, nparts, K, post, param, out);
  return cudaGetLastError();
}

// a *= (alpha from device memory: coef[0] + i coef[1])
__global__ void __launch_bounds__(RED_THREADS)
k_scale_dev(double2* __restrict__ state, uint64_t count, const double* __restrict__ coef) {
  const double2 al{coef[0], coef[1]};
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  for (; i + 3 * stride < count; i += 4 * stride) {
    const double2 a0 = __ldcs(state + i), a1 = __ldcs(state + i + stride), a2 = __ldcs(state + i + 2 * stride), a3 = __ldcs(state + i + 3 * stride);
    __stcs(state + i, cmul(al, a0)); __stcs(state + i + stride, cmul(al, a1));
    __stcs(state + i + 2 * stride, cmul(al, a2)); __stcs(state + i + 3 * stride, cmul(al, a3));
  }
  for (; i < count; i += stride) state[i] = cmul(al, state[i]);
}
cudaError_t launch_scale_dev(double2* state, uint64_t count, const double* coef, int grid, cudaStream_t s) {
  k_scale_dev<<<grid, RED_THREADS, 0, s>>>(state, count, coef);
  return cudaGetLastError();
}

// One Grover iteration as a single streaming pass (32 B per amplitude): a' = alpha * a + beta with (alpha, beta) =
// (-1, 2*mean) left in coef[0..4) by the preceding sum, then the sign flips of the phase oracles that follow the diffusion
// (marked = local indices on this rank), then - when another diffusion follows - the sum of the new amplitudes for it
// (partials: [grid][2], finalised by k_finalize exactly like k_reduce's).
__global__ void __launch_bounds__(RED_THREADS)
k_grover_step(double2* __restrict__ state, uint64_t count, const double* __restrict__ coef, GroverMarks marks,
              double* __restrict__ partials) {
  __shared__ double sm[2 * 32];
  const double2 al{coef[0], coef[1]}, be{coef[2], coef[3]};
  double v[2] = {0.0, 0.0};
  auto step = [&](uint64_t i, double2 a) {
    double2 t = cmul(al, a);
    t.x += be.x; t.y += be.y;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < marks.n && i == marks.idx[k]) { t.x = -t.x; t.y = -t.y; }
    __stcs(state + i, t);
    v[0] += t.x; v[1] += t.y;
  };
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  // four independent 16-byte loads in flight per thread before the first store (a store to `state` would otherwise
  // order the next iteration's load behind it)
  for (; i + 3 * stride < count; i += 4 * stride) {
    const double2 a0 = __ldcs(state + i), a1 = __ldcs(state + i + stride), a2 = __ldcs(state + i + 2 * stride),
                  a3 = __ldcs(state + i + 3 * stride);
    step(i, a0); step(i + stride, a1); step(i + 2 * stride, a2); step(i + 3 * stride, a3);
  }
  for (; i < count; i += stride) step(i, __ldcs(state + i));
  if (partials) {
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = v[0]; partials[2 * blockIdx.x + 1] = v[1]; }
  }
}
cudaError_t launch_grover_step(double2* state, uint64_t count, const double* coef, const GroverMarks& marks, double* partials,
                               int grid, cudaStream_t s) {
  k_grover_step<<<grid, RED_THREADS, 0, s>>>(state, count, coef, marks, partials);
  return cudaGetLastError();
}

// p_i = |a_i|^2 (domain/state.clj:676-682)
__global__ void __launch_bounds__(RED_THREADS)
k_probabilities(const double2* __restrict__ state, uint64_t offset, uint64_t count, double* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  const double2* src = state + offset;
  uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  for (; i + 3 * stride < count; i += 4 * stride) {
    const double2 a0 = __ldcs(src + i), a1 = __ldcs(src + i + stride), a2 = __ldcs(src + i + 2 * stride), a3 = __ldcs(src + i + 3 * stride);
    out[i] = a0.x * a0.x + a0.y * a0.y; out[i + stride] = a1.x * a1.x + a1.y * a1.y;
    out[i + 2 * stride] = a2.x * a2.x + a2.y * a2.y; out[i + 3 * stride] = a3.x * a3.x + a3.y * a3.y;
  }
  for (; i < count; i += stride) { const double2 a = src[i]; out[i] = a.x * a.x + a.y * a.y; }
}
cudaError_t launch_probabilities(const double2* state, uint64_t offset, uint64_t count, double* out, int grid, cudaStream_t s) {
  k_probabilities<<<grid, RED_THREADS, 0, s>>>(state, offset, count, out);
  return cudaGetLastError();
}

__global__ void k_gather(const double2* __restrict__ state, const uint64_t* __restrict__ idx, uint64_t n, double2* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = state[idx[i]];
}
cudaError_t launch_gather(const double2* state, const uint64_t* idx, uint64_t n, double2* out, cudaStream_t s) {
  k_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(state, idx, n, out);
  return cudaGetLastError();
}

// <phi|psi> = sum conj(phi_i) psi_i ; partials [grid][2]
__global__ void __launch_bounds__(RED_THREADS)
k_inner(const double2* __restrict__ phi, const double2* __restrict__ psi, uint64_t count, double* __restrict__ partials) {
  __shared__ double sm[2 * 32];
  double v[2] = {0.0, 0.0};
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  for (; i + stride < count; i += 2 * stride) {
    const double2 a0 = __ldcs(phi + i), b0 = __ldcs(psi + i), a1 = __ldcs(phi + i + stride), b1 = __ldcs(psi + i + stride);
    v[0] += (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y);
    v[1] += (a0.x * b0.y - a0.y * b0.x) + (a1.x * b1.y - a1.y * b1.x);
  }
  for (; i < count; i += stride) {
    const double2 a = phi[i], b = psi[i];
    v[0] += a.x * b.x + a.y * b.y;
    v[1] += a.x * b.y - a.y * b.x;
  }
  block_sum<2>(v, sm);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = v[0]; partials[2 * blockIdx.x + 1] = v[1]; }
}
cudaError_t launch_inner(const double2* phi, const double2* psi, uint64_t count, double* partials, int grid, cudaStream_t s) {
  k_inner<<<grid, RED_THREADS, 0, s>>>(phi, psi, count, partials);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ Pauli expectation (x-mask / z-mask form)
// Terms of one group share the X mask.  For a pair (i, j = i ^ xmask), i < j (bit `pivot` of i is 0):
//   Re(conj(a_i) c_j a_j + conj(a_j) c_i a_i),  c_k = phase * (-1)^popc(k & zmask), phase = i^ny.
// For xmask = 0 every index is its own partner: contribution (-1)^popc(i & z) |a_i|^2.
// ext_or supplies the rank bits of the global index.  partials: [grid][EXPECT_TERMS]
__global__ void __launch_bounds__(RED_THREADS)
k_expect_group(const double2* __restrict__ state, uint64_t count, uint64_t xmask, int pivot, uint64_t ext_or,
               ExpectTerms terms, double* __restrict__ partials) {
  __shared__ double sm[EXPECT_TERMS * 32];
  double acc[EXPECT_TERMS];
#pragma unroll
  for (int k = 0; k < EXPECT_TERMS; ++k) acc[k] = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  if (xmask == 0 && count >= 2048 && (count & 2047) == 0) {
    // Diagonal terms, large states: sum_i (-1)^parity(i & z) |a_i|^2 without a POPC per term and amplitude (POPC issues at a
    // quarter of the integer rate and bounded this loop at 0.58 - 0.63 of the HBM peak).  A warp walks blocks of 2^11
    // amplitudes; in step t lane l takes the four amplitudes t << 7 | j << 5 | l, j = 0..3 (every load instruction of the warp
    // reads 512 contiguous bytes), so parity(index & z) = parity(base & z) ^ parity(t << 7 & z) ^ parity(l & z) ^ (z's bits 5
    // and 6 against j, folded into the choice of p0 +- p1 +- p2 +- p3).  The 16 terms' parities travel as the bits of one
    // word: base part once per block, lane part once per kernel, step part from a 16-entry table.
    __shared__ uint32_t steppar[16];
    if (threadIdx.x < 16) {
      uint32_t w = 0;
      for (int k = 0; k < EXPECT_TERMS; ++k) w |= (uint32_t)(__popcll(((uint64_t)threadIdx.x << 7) & terms.zmask[k]) & 1) << k;
      steppar[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t lanepar = 0;
    for (int k = 0; k < EXPECT_TERMS; ++k) lanepar |= (uint32_t)(__popcll((uint64_t)lane & terms.zmask[k]) & 1) << k;
    const uint64_t nblk = count >> 11, wstride = (uint64_t)gridDim.x * (RED_THREADS / 32);
    for (uint64_t blk = (uint64_t)blockIdx.x * (RED_THREADS / 32) + (threadIdx.x >> 5); blk < nblk; blk += wstride) {
      const uint64_t base = blk << 11;
      uint32_t par0 = lanepar;
#pragma unroll
      for (int k = 0; k < EXPECT_TERMS; ++k) par0 ^= (uint32_t)(__popcll((base | ext_or) & terms.zmask[k]) & 1) << k;
      const double2* p = state + base + lane;
      double2 a0 = __ldcs(p), a1 = __ldcs(p + 32), a2 = __ldcs(p + 64), a3 = __ldcs(p + 96);
#pragma unroll 4
      for (uint32_t t = 0; t < 16; ++t) {
        double2 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
        if (t + 1 < 16) { const double2* pn = p + ((uint64_t)(t + 1) << 7); b0 = __ldcs(pn); b1 = __ldcs(pn + 32); b2 = __ldcs(pn + 64); b3 = __ldcs(pn + 96); }
        const double p0 = a0.x * a0.x + a0.y * a0.y, p1 = a1.x * a1.x + a1.y * a1.y, p2 = a2.x * a2.x + a2.y * a2.y, p3 = a3.x * a3.x + a3.y * a3.y;
        // combo[c]: bit 0 of c = z bit 5 set (p1, p3 flip), bit 1 = z bit 6 set (p2, p3 flip)
        const double combo[4] = {(p0 + p1) + (p2 + p3), (p0 - p1) + (p2 - p3), (p0 + p1) - (p2 + p3), (p0 - p1) - (p2 - p3)};
        const uint32_t w = par0 ^ steppar[t];
#pragma unroll
        for (int k = 0; k < EXPECT_TERMS; ++k)
          if (k < terms.n) {
            const double v = combo[((uint32_t)terms.zmask[k] >> 5) & 3u];
            const int vh = __double2hiint(v) ^ (int)((w << (31 - k)) & 0x80000000u);
            acc[k] += __hiloint2double(vh, __double2loint(v));
          }
        a0 = b0; a1 = b1; a2 = b2; a3 = b3;
      }
    }
  } else if (xmask == 0) {
    // Diagonal terms: sum_i (-1)^parity(i & z) |a_i|^2.  A thread takes FOUR consecutive amplitudes (64 bytes): their
    // parities differ from the first one's only through z's two lowest bits, which are constants of the term, so one
    // parity evaluation (two ANDs, one XOR, ONE 32-bit POPC - POPC issues at a quarter of the integer rate and bounded the
    // one-amplitude-per-evaluation loop at 1.3 TB/s, profiles/r2d_reductions.txt) signs one of four precomputed
    // combinations p0 +- p1 +- p2 +- p3.  The four loads of the NEXT quad are issued before the arithmetic of the current
    // one.  (Choosing the combination with selects and signing it with integer instructions instead of the indexed local
    // array + negate was measured slower: 4.87 vs 4.56 ms at 30 qubits, profiles/r2r_expect_*.log.)
    const uint64_t quads = count >> 2;
    uint64_t q = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
    double2 a0, a1, a2, a3;
    if (q < quads) { const double2* p = state + (q << 2); a0 = __ldcs(p); a1 = __ldcs(p + 1); a2 = __ldcs(p + 2); a3 = __ldcs(p + 3); }
    while (q < quads) {
      const uint64_t qn = q + stride;
      double2 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
      if (qn < quads) { const double2* p = state + (qn << 2); b0 = __ldcs(p); b1 = __ldcs(p + 1); b2 = __ldcs(p + 2); b3 = __ldcs(p + 3); }
      const double p0 = a0.x * a0.x + a0.y * a0.y, p1 = a1.x * a1.x + a1.y * a1.y, p2 = a2.x * a2.x + a2.y * a2.y, p3 = a3.x * a3.x + a3.y * a3.y;
      // combo[c]: bit 0 of c = z bit 0 set (p1, p3 flip), bit 1 = z bit 1 set (p2, p3 flip)
      const double combo[4] = {(p0 + p1) + (p2 + p3), (p0 - p1) + (p2 - p3), (p0 + p1) - (p2 + p3), (p0 - p1) - (p2 - p3)};
      const uint64_t gi = (q << 2) | ext_or;
      const uint32_t lo = (uint32_t)gi, hi = (uint32_t)(gi >> 32);
#pragma unroll
      for (int k = 0; k < EXPECT_TERMS; ++k)
        if (k < terms.n) {
          const uint32_t zl = (uint32_t)terms.zmask[k];
          const uint32_t f = (lo & zl) ^ (hi & (uint32_t)(terms.zmask[k] >> 32));        // lo's bits 0, 1 are zero
          const double v = combo[zl & 3u];
          acc[k] += (__popc(f) & 1) ? -v : v;
        }
      a0 = b0; a1 = b1; a2 = b2; a3 = b3;
      q = qn;
    }
    for (uint64_t i = (quads << 2) + (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < count; i += stride) {   // count < 4
      const double2 a = state[i];
      const double p = a.x * a.x + a.y * a.y;
      const uint64_t gi = i | ext_or;
#pragma unroll
      for (int k = 0; k < EXPECT_TERMS; ++k)
        if (k < terms.n) acc[k] += (__popcll(gi & terms.zmask[k]) & 1) ? -p : p;
    }
  } else {
    const uint64_t half = count >> 1;
    auto pair_index = [&](uint64_t h) { return ((h >> pivot) << (pivot + 1)) | (h & ((1ULL << pivot) - 1ULL)); };
    auto one = [&](uint64_t i, double2 ai, double2 aj) {
      const uint64_t j = i ^ xmask;
      // conj(ai)*aj = (x, y);  conj(aj)*ai = (x, -y)
      const double x = ai.x * aj.x + ai.y * aj.y, y = ai.x * aj.y - ai.y * aj.x;
      const uint64_t gi = i | ext_or, gj = j | ext_or;
#pragma unroll
      for (int k = 0; k < EXPECT_TERMS; ++k) {
        if (k < terms.n) {
          const uint32_t zl = (uint32_t)terms.zmask[k], zh = (uint32_t)(terms.zmask[k] >> 32);
          const uint32_t fi = __popc(((uint32_t)gi & zl) ^ ((uint32_t)(gi >> 32) & zh)) & 1u;
          const uint32_t fj = __popc(((uint32_t)gj & zl) ^ ((uint32_t)(gj >> 32) & zh)) & 1u;
          // Re(ph * sj * (x + iy)) + Re(ph * si * (x - iy)) with ph = (pr, pi), si / sj = +-1:
          //   = (sj + si) pr x - (sj - si) pi y: equal signs leave the x part, opposite signs the y part
          const double u = terms.pr[k] * x, w = terms.pi[k] * y;
          const double v = (fi == fj) ? 2.0 * u : -2.0 * w;
          const int vh = __double2hiint(v) ^ (int)(fj << 31);
          acc[k] += __hiloint2double(vh, __double2loint(v));
        }
      }
    };
    // two pairs = four 16-byte loads per step; the loads of the next step are issued before the arithmetic of this one
    uint64_t h = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
    double2 a0, b0, a1, b1;
    uint64_t i0 = 0, i1 = 0;
    bool have = h + stride < half;
    if (have) { i0 = pair_index(h); i1 = pair_index(h + stride); a0 = __ldcs(state + i0); b0 = __ldcs(state + (i0 ^ xmask)); a1 = __ldcs(state + i1); b1 = __ldcs(state + (i1 ^ xmask)); }
    while (have) {
      const uint64_t hn = h + 2 * stride;
      const bool have_n = hn + stride < half;
      double2 na0 = a0, nb0 = b0, na1 = a1, nb1 = b1;
      uint64_t ni0 = 0, ni1 = 0;
      if (have_n) { ni0 = pair_index(hn); ni1 = pair_index(hn + stride); na0 = __ldcs(state + ni0); nb0 = __ldcs(state + (ni0 ^ xmask)); na1 = __ldcs(state + ni1); nb1 = __ldcs(state + (ni1 ^ xmask)); }
      one(i0, a0, b0); one(i1, a1, b1);
      a0 = na0; b0 = nb0; a1 = na1; b1 = nb1; i0 = ni0; i1 = ni1;
      h = hn; have = have_n;
    }
    for (; h < half; h += stride) { const uint64_t i = pair_index(h); one(i, state[i], state[i ^ xmask]); }
  }
  block_sum<EXPECT_TERMS>(acc, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < EXPECT_TERMS; ++k) partials[(uint64_t)blockIdx.x * EXPECT_TERMS + k] = acc[k];
  }
}
cudaError_t launch_expect_group(const double2* state, uint64_t count, uint64_t xmask, int pivot, uint64_t ext_or,
                                const ExpectTerms& terms, double* partials, int grid, cudaStream_t s) {
  k_expect_group<<<grid, RED_THREADS, 0, s>>>(state, count, xmask, pivot, ext_or, terms, partials);
  return cudaGetLastError();
}

// <psi| (I..O_t..I) |psi> for a 2x2 observable on index bit `bit`; partials [grid][2] (re, im)
__global__ void __launch_bounds__(RED_THREADS)
k_expect_1q(const double2* __restrict__ state, uint64_t count, int bit, Mat2 O, double* __restrict__ partials) {
  __shared__ double sm[2 * 32];
  double v[2] = {0.0, 0.0};
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS, half = count >> 1;
  auto one = [&](double2 a0, double2 a1) {
    const double2 b0 = cmul2(double2{O.m[0], O.m[1]}, a0, double2{O.m[2], O.m[3]}, a1);
    const double2 b1 = cmul2(double2{O.m[4], O.m[5]}, a0, double2{O.m[6], O.m[7]}, a1);
    v[0] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
    v[1] += a0.x * b0.y - a0.y * b0.x + a1.x * b1.y - a1.y * b1.x;
  };
  auto idx0 = [&](uint64_t h) { return ((h >> bit) << (bit + 1)) | (h & ((1ULL << bit) - 1ULL)); };
  uint64_t h = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  for (; h + stride < half; h += 2 * stride) {          // two pairs = four 16-byte loads in flight
    const uint64_t i0 = idx0(h), j0 = idx0(h + stride);
    const double2 a0 = __ldcs(state + i0), a1 = __ldcs(state + (i0 | (1ULL << bit))), c0 = __ldcs(state + j0), c1 = __ldcs(state + (j0 | (1ULL << bit)));
    one(a0, a1); one(c0, c1);
  }
  for (; h < half; h += stride) { const uint64_t i0 = idx0(h); one(state[i0], state[i0 | (1ULL << bit)]); }
  block_sum<2>(v, sm);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = v[0]; partials[2 * blockIdx.x + 1] = v[1]; }
}
cudaError_t launch_expect_1q(const double2* state, uint64_t count, int bit, const Mat2& O, double* partials, int grid, cudaStream_t s) {
  k_expect_1q<<<grid, RED_THREADS, 0, s>>>(state, count, bit, O, partials);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ partial measurement (domain/state.clj:946-1014)
// key(i) = sum_k bit(i, pos[k]) << k ; histogram of |a|^2 over keys, per block in shared memory.
// partials: [grid][2^m]
// Deterministic (no floating-point atomics): every warp owns a private histogram in shared memory.  A lane's index is
// (multiple of 32) + lane, so inside a warp the key varies only with the measured (or filtered) bits among index bits
// 0..4: the lanes that share a key differ exactly in the other low bits, and their values are combined with xor-shuffles
// over those bits (same additions in the same order in every lane of the group); the lane with zeros there alone updates
// the bin - no two lanes ever write one bin, every addition has a fixed order, and the warps' histograms are added in warp
// order.  blockDim = 32 * W with W * 2^m * 8 bytes <= 64 KB (W = 8 up to m = 10, 2 at m = 12).  Optional filter (fl.n > 0):
// only amplitudes whose bits fl.pos match fval are counted (second pass of a measurement of more than 12 qubits).
__global__ void __launch_bounds__(RED_THREADS)
k_marginal(const double2* __restrict__ state, uint64_t count, uint64_t ext_or, BitList bl, BitList fl, uint32_t fval,
           double* __restrict__ partials) {
  extern __shared__ double hist[];
  const uint32_t nk = 1u << bl.n, nthreads = blockDim.x, nw = nthreads >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t k = threadIdx.x; k < nk * nw; k += nthreads) hist[k] = 0.0;
  __syncthreads();
  double* mine = hist + (size_t)wid * nk;
  uint32_t fixed_low = 0;                                          // index bits 0..4 that the key or the filter depends on
  for (int k = 0; k < bl.n; ++k) if (bl.pos[k] < 5) fixed_low |= 1u << bl.pos[k];
  for (int k = 0; k < fl.n; ++k) if (fl.pos[k] < 5) fixed_low |= 1u << fl.pos[k];
  const uint32_t free_low = ~fixed_low & 31u;
  const bool leader = (lane & free_low) == 0;
  const uint64_t stride = (uint64_t)gridDim.x * nthreads;
  const uint64_t rounds = (count + stride - 1) / stride;           // every lane runs the same number of iterations (warp collectives)
  // one element: key / validity / probability, combined over the group's lanes, added to the warp's bin by the leader
  auto consume = [&](uint64_t i, double2 a) {
    uint32_t key = 0;
    bool valid = false;
    double p = 0.0;
    if (i < count) {
      const uint64_t gi = i | ext_or;
      uint32_t f = 0;
      for (int k = 0; k < fl.n; ++k) f |= (uint32_t)((gi >> fl.pos[k]) & 1ULL) << k;
      if (f == fval) {
        valid = true;
#pragma unroll 4
        for (int k = 0; k < bl.n; ++k) key |= (uint32_t)((gi >> bl.pos[k]) & 1ULL) << k;
        p = a.x * a.x + a.y * a.y;
      }
    }
#pragma unroll
    for (uint32_t b = 0; b < 5; ++b)
      if ((free_low >> b) & 1u) p += __shfl_xor_sync(0xffffffffu, p, 1u << b);
    // a group lies either wholly inside or wholly outside the range / the filter, except in the ragged last iteration,
    // where the lanes beyond `count` contribute zeros
    if (leader && valid) mine[key] += p;
    __syncwarp();
  };
  const uint64_t first = (uint64_t)blockIdx.x * nthreads + threadIdx.x;
  uint64_t it = 0;
  for (; it + 3 < rounds; it += 4) {                               // four independent 16-byte loads in flight per thread
    const uint64_t i0 = it * stride + first, i1 = i0 + stride, i2 = i1 + stride, i3 = i2 + stride;
    const double2 z{0.0, 0.0};
    const double2 a0 = i0 < count ? __ldcs(state + i0) : z, a1 = i1 < count ? __ldcs(state + i1) : z;
    const double2 a2 = i2 < count ? __ldcs(state + i2) : z, a3 = i3 < count ? __ldcs(state + i3) : z;
    consume(i0, a0); consume(i1, a1); consume(i2, a2); consume(i3, a3);
  }
  for (; it < rounds; ++it) {
    const uint64_t i = it * stride + first;
    consume(i, i < count ? __ldcs(state + i) : double2{0.0, 0.0});
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nk; k += nthreads) {
    double t = 0.0;
    for (uint32_t w = 0; w < nw; ++w) t += hist[(size_t)w * nk + k];
    partials[(uint64_t)blockIdx.x * nk + k] = t;
  }
}
// The same histogram for large states, block-wise: six index bits >= 5 that neither the key nor the filter depends on (freeb)
// become a lane's private loop, so a lane adds up its 64 amplitudes in a register and the warp combines and bins ONCE per block
// of 2^11 amplitudes instead of once per amplitude (the generic kernel spends ten SHFL per 16 bytes when no measured bit lies
// in the lane bits).  Every load instruction of a warp reads 512 contiguous bytes; blocks the filter rejects are not read.
// base_pos: the positions of the remaining index bits (block number -> base address).  Deterministic like k_marginal.
struct BasePos { int n; int pos[40]; };
__global__ void __launch_bounds__(RED_THREADS)
k_marginal_blocks(const double2* __restrict__ state, uint64_t nblk, uint64_t ext_or, BitList bl, BitList fl, uint32_t fval, BitList freeb,
                  BasePos bp, double* __restrict__ partials) {
  extern __shared__ double hist[];
  __shared__ uint64_t offs[64];
  const uint32_t nk = 1u << bl.n, nthreads = blockDim.x, nw = nthreads >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t k = threadIdx.x; k < nk * nw; k += nthreads) hist[k] = 0.0;
  for (uint32_t i = threadIdx.x; i < 64; i += nthreads) {
    uint64_t o = 0;
    for (int b = 0; b < 6; ++b) o |= (uint64_t)((i >> b) & 1u) << freeb.pos[b];
    offs[i] = o;
  }
  __syncthreads();
  double* mine = hist + (size_t)wid * nk;
  uint32_t fixed_low = 0;                                          // lane bits that the key or the filter depends on
  for (int k = 0; k < bl.n; ++k) if (bl.pos[k] < 5) fixed_low |= 1u << bl.pos[k];
  for (int k = 0; k < fl.n; ++k) if (fl.pos[k] < 5) fixed_low |= 1u << fl.pos[k];
  const uint32_t free_low = ~fixed_low & 31u;
  const bool leader = (lane & free_low) == 0;
  const uint64_t wstride = (uint64_t)gridDim.x * nw;
  for (uint64_t blk = (uint64_t)blockIdx.x * nw + wid; blk < nblk; blk += wstride) {
    uint64_t base = 0;
    for (int b = 0; b < bp.n; ++b) base |= ((blk >> b) & 1ULL) << bp.pos[b];
    const uint64_t gi = base | lane | ext_or;
    uint32_t f = 0, key = 0;
    for (int k = 0; k < fl.n; ++k) f |= (uint32_t)((gi >> fl.pos[k]) & 1ULL) << k;
    const bool valid = f == fval;
    if (!__any_sync(0xffffffffu, valid)) continue;                 // the whole block fails the filter: not read at all
#pragma unroll 4
    for (int k = 0; k < bl.n; ++k) key |= (uint32_t)((gi >> bl.pos[k]) & 1ULL) << k;
    const double2* p = state + base + lane;
    double sum = 0.0;
    double2 a0 = __ldcs(p + offs[0]), a1 = __ldcs(p + offs[1]), a2 = __ldcs(p + offs[2]), a3 = __ldcs(p + offs[3]);
#pragma unroll 2
    for (uint32_t t = 0; t < 64; t += 4) {
      double2 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
      if (t + 4 < 64) { b0 = __ldcs(p + offs[t + 4]); b1 = __ldcs(p + offs[t + 5]); b2 = __ldcs(p + offs[t + 6]); b3 = __ldcs(p + offs[t + 7]); }
      sum += (a0.x * a0.x + a0.y * a0.y) + (a1.x * a1.x + a1.y * a1.y);
      sum += (a2.x * a2.x + a2.y * a2.y) + (a3.x * a3.x + a3.y * a3.y);
      a0 = b0; a1 = b1; a2 = b2; a3 = b3;
    }
    if (!valid) sum = 0.0;
#pragma unroll
    for (uint32_t b = 0; b < 5; ++b)
      if ((free_low >> b) & 1u) sum += __shfl_xor_sync(0xffffffffu, sum, 1u << b);
    if (leader && valid) mine[key] += sum;
    __syncwarp();
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nk; k += nthreads) {
    double t = 0.0;
    for (uint32_t w = 0; w < nw; ++w) t += hist[(size_t)w * nk + k];
    partials[(uint64_t)blockIdx.x * nk + k] = t;
  }
}

cudaError_t launch_marginal(const double2* state, uint64_t count, uint64_t ext_or, const BitList& bl, const BitList& fl, uint32_t fval,
                            double* partials, int grid, cudaStream_t s) {
  int nw = (int)(8192u >> bl.n);
  nw = nw < 1 ? 1 : (nw > 8 ? 8 : nw);
  // block-wise kernel: power-of-two slice of >= 2^11 amplitudes with six free index bits >= 5
  if (count >= 2048 && (count & (count - 1)) == 0) {
    int nl = 0;
    while ((1ULL << nl) < count) ++nl;
    uint64_t used = 31;                                           // lane bits
    for (int k = 0; k < bl.n; ++k) if (bl.pos[k] < nl) used |= 1ULL << bl.pos[k];
    for (int k = 0; k < fl.n; ++k) if (fl.pos[k] < nl) used |= 1ULL << fl.pos[k];
    BitList freeb; freeb.n = 0;
    for (int p = 5; p < nl && freeb.n < 6; ++p) if (!((used >> p) & 1)) { freeb.pos[freeb.n++] = p; used |= 1ULL << p; }
    if (freeb.n == 6) {
      BasePos bp; bp.n = 0;
      uint64_t taken = 31;
      for (int k = 0; k < 6; ++k) taken |= 1ULL << freeb.pos[k];
      for (int p = 5; p < nl; ++p) if (!((taken >> p) & 1)) bp.pos[bp.n++] = p;
      static std::atomic<uint64_t> configured_b{0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!(configured_b.load() & (1ULL << (dev & 63)))) {
        cudaError_t e = cudaFuncSetAttribute(k_marginal_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        configured_b.fetch_or(1ULL << (dev & 63));
      }
      k_marginal_blocks<<<grid, 32 * nw, (sizeof(double) << bl.n) * nw, s>>>(state, count >> 11, ext_or, bl, fl, fval, freeb, bp, partials);
      return cudaGetLastError();
    }
  }
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(configured.load() & (1ULL << (dev & 63)))) {
    cudaError_t e = cudaFuncSetAttribute(k_marginal, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    configured.fetch_or(1ULL << (dev & 63));
  }
  k_marginal<<<grid, 32 * nw, (sizeof(double) << bl.n) * nw, s>>>(state, count, ext_or, bl, fl, fval, partials);
  return cudaGetLastError();
}

// collapse: keep amplitudes whose key == sel (scaled by factor), zero the rest
__global__ void __launch_bounds__(RED_THREADS)
k_collapse(double2* __restrict__ state, uint64_t count, uint64_t ext_or, BitList bl, uint32_t sel, double factor) {
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  for (uint64_t i = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < count; i += stride) {
    const uint64_t gi = i | ext_or;
    uint32_t key = 0;
#pragma unroll 4
    for (int k = 0; k < bl.n; ++k) key |= (uint32_t)((gi >> bl.pos[k]) & 1ULL) << k;
    if (key == sel) { double2 a = state[i]; a.x *= factor; a.y *= factor; state[i] = a; }
    else state[i] = double2{0.0, 0.0};                       // no read needed for the amplitudes that vanish
  }
}
cudaError_t launch_collapse(double2* state, uint64_t count, uint64_t ext_or, const BitList& bl, uint32_t sel, double factor, int grid, cudaStream_t s) {
  k_collapse<<<grid, RED_THREADS, 0, s>>>(state, count, ext_or, bl, sel, factor);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ sampling (domain/state.clj:894-913)
// In-block inclusive scan of SAMPLE_CHUNK probabilities (SAMPLE_THREADS threads x SAMPLE_PER_THREAD each).
// Returns this thread's inclusive prefix of its last element and fills `loc[]` with per-element
// inclusive prefixes inside the chunk.  Both k_chunk_sums and k_sample use exactly this code so that
// the chunk total and the in-chunk scan round identically.
__device__ __forceinline__ double chunk_scan(const double2* __restrict__ src, uint64_t valid, double (&loc)[SAMPLE_PER_THREAD],
                                             double* sm, double* smp) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // phase A: coalesced loads, |a|^2 into padded shared memory (index i + i/16: conflict-free in phase B)
#pragma unroll
  for (int k = 0; k < SAMPLE_PER_THREAD; ++k) {
    const uint32_t i = (uint32_t)k * SAMPLE_THREADS + tid;
    double p = 0.0;
    if (i < valid) { const double2 a = src[i]; p = a.x * a.x + a.y * a.y; }
    smp[i + (i >> 4)] = p;
  }
  __syncthreads();
  // phase B: each thread owns SAMPLE_PER_THREAD consecutive elements
  double run = 0.0;
#pragma unroll
  for (int k = 0; k < SAMPLE_PER_THREAD; ++k) {
    run += smp[tid * (SAMPLE_PER_THREAD + 1) + k];
    loc[k] = run;
  }
  // exclusive scan of per-thread totals across the block
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) sm[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    double w = (lane < SAMPLE_THREADS / 32) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    sm[32 + lane] = w;   // inclusive prefix of warp totals
  }
  __syncthreads();
  const double warp_excl = (wid == 0) ? 0.0 : sm[32 + wid - 1];
  const double thread_excl = warp_excl + (incl - run);
#pragma unroll
  for (int k = 0; k < SAMPLE_PER_THREAD; ++k) loc[k] += thread_excl;
  const double total = sm[32 + SAMPLE_THREADS / 32 - 1];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
k_chunk_sums(const double2* __restrict__ state, uint64_t count, double* __restrict__ sums) {
  __shared__ double sm[64];
  __shared__ double smp[SAMPLE_CHUNK + SAMPLE_CHUNK / 16];
  double loc[SAMPLE_PER_THREAD];
  for (uint64_t c = blockIdx.x; c * SAMPLE_CHUNK < count; c += gridDim.x) {
    const uint64_t start = c * SAMPLE_CHUNK;
    const uint64_t valid = (count - start < SAMPLE_CHUNK) ? (count - start) : SAMPLE_CHUNK;
    const double total = chunk_scan(state + start, valid, loc, sm, smp);
    if (threadIdx.x == 0) sums[c] = total;
  }
}

// in-place inclusive scan of `n` doubles by one block: tiles of RED_THREADS x SCAN_ITEMS elements, each thread scans its
// SCAN_ITEMS consecutive elements in registers, the thread totals are combined with warp shuffles, a carry links the tiles
// (2^30 amplitudes = 262 144 chunk sums = 64 tiles; one element per thread and tile made this 1024 serial steps)
constexpr int SCAN_ITEMS = 16;
__global__ void __launch_bounds__(RED_THREADS)
k_scan_inclusive(double* __restrict__ v, uint64_t n) {
  __shared__ double sm[64];
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double carry = 0.0;                                        // inclusive prefix of everything before this tile (same in every thread)
  for (uint64_t base = 0; base < n; base += (uint64_t)RED_THREADS * SCAN_ITEMS) {
    const uint64_t i0 = base + (uint64_t)tid * SCAN_ITEMS;
    double x[SCAN_ITEMS];
    if (i0 + SCAN_ITEMS <= n) {
      const double2* src = reinterpret_cast<const double2*>(v + i0);   // v is 256-byte aligned, i0 a multiple of 16
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS / 2; ++k) { const double2 t = src[k]; x[2 * k] = t.x; x[2 * k + 1] = t.y; }
    } else {
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS; ++k) x[k] = (i0 + k < n) ? v[i0 + k] : 0.0;
    }
#pragma unroll
    for (int k = 1; k < SCAN_ITEMS; ++k) x[k] += x[k - 1];
    double incl = x[SCAN_ITEMS - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      double w = (lane < RED_THREADS / 32) ? sm[lane] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      sm[32 + lane] = w;
    }
    __syncthreads();
    const double excl = carry + ((wid == 0) ? 0.0 : sm[32 + wid - 1]) + (incl - x[SCAN_ITEMS - 1]);
    carry += sm[32 + RED_THREADS / 32 - 1];
    if (i0 + SCAN_ITEMS <= n) {
      double2* dst = reinterpret_cast<double2*>(v + i0);
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS / 2; ++k) dst[k] = double2{excl + x[2 * k], excl + x[2 * k + 1]};
    } else {
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS; ++k) if (i0 + k < n) v[i0 + k] = excl + x[k];
    }
    __syncthreads();                                         // sm[] is rewritten by the next tile
  }
}

// One block per shot.  r = total * u (+ nothing); `cum_chunks` = inclusive prefix of chunk sums, offset by
// `rank_offset` (probability mass of lower ranks).  outcome = first index with cum >= r, clamped.
// Shots whose r lies outside this rank's (lo, hi] range are left untouched (multi-GPU).
__global__ void __launch_bounds__(SAMPLE_THREADS)
k_sample(const double2* __restrict__ state, uint64_t count, const double* __restrict__ cum_chunks, uint64_t n_chunks,
         const double* __restrict__ uniforms, uint64_t n_shots, double total, double rank_offset, int is_first_rank,
         int is_last_rank, uint64_t index_offset, unsigned long long* __restrict__ outcomes) {
  __shared__ double sm[64];
  __shared__ double smp[SAMPLE_CHUNK + SAMPLE_CHUNK / 16];
  __shared__ unsigned long long best;
  double loc[SAMPLE_PER_THREAD];
  for (uint64_t shot = blockIdx.x; shot < n_shots; shot += gridDim.x) {
    const double r = total * uniforms[shot];
    const double local_total = cum_chunks[n_chunks - 1];
    const double rl = r - rank_offset;                       // position inside this rank's mass
    // ownership: first rank takes rl <= local_total (incl. r <= 0); others (0, local_total]; last rank also r > all
    const bool below = !(rl > 0.0) && !is_first_rank;
    const bool above = (rl > local_total) && !is_last_rank;
    if (below || above) continue;                           // uniform across the block
    // lower_bound over chunks: first c with cum_chunks[c] >= rl
    uint64_t lo = 0, hi = n_chunks;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (cum_chunks[mid] < rl) lo = mid + 1; else hi = mid; }
    uint64_t c = (lo < n_chunks) ? lo : n_chunks - 1;
    const double excl = (c == 0) ? 0.0 : cum_chunks[c - 1];
    const uint64_t start = c * SAMPLE_CHUNK;
    const uint64_t valid = (count - start < SAMPLE_CHUNK) ? (count - start) : SAMPLE_CHUNK;
    if (threadIdx.x == 0) best = ~0ULL;
    chunk_scan(state + start, valid, loc, sm, smp);         // has __syncthreads inside
    unsigned long long mine = ~0ULL;
#pragma unroll
    for (int k = SAMPLE_PER_THREAD - 1; k >= 0; --k) {
      const uint64_t i = (uint64_t)threadIdx.x * SAMPLE_PER_THREAD + k;
      if (i < valid && !(excl + loc[k] < rl)) mine = i;
    }
    if (mine != ~0ULL) atomicMin(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t o = (best == ~0ULL) ? (start + valid - 1) : (start + best);
      outcomes[shot] = o + index_offset;
    }
    __syncthreads();
  }
}

cudaError_t launch_chunk_sums(const double2* state, uint64_t count, double* sums, int grid, cudaStream_t s) {
  k_chunk_sums<<<grid, SAMPLE_THREADS, 0, s>>>(state, count, sums);
  return cudaGetLastError();
}
cudaError_t launch_scan_inclusive(double* v, uint64_t n, cudaStream_t s) {
  k_scan_inclusive<<<1, RED_THREADS, 0, s>>>(v, n);
  return cudaGetLastError();
}
cudaError_t launch_sample(const double2* state, uint64_t count, const double* cum_chunks, uint64_t n_chunks, const double* uniforms,
                          uint64_t n_shots, double total, double rank_offset, int first, int last, uint64_t index_offset,
                          unsigned long long* outcomes, int grid, cudaStream_t s) {
  k_sample<<<grid, SAMPLE_THREADS, 0, s>>>(state, count, cum_chunks, n_chunks, uniforms, n_shots, total, rank_offset, first, last,
                                           index_offset, outcomes);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ chunked half exchange helpers (multi-GPU)
// pack / unpack the half of the local slice whose bit `lbit` equals `want` into / from a contiguous buffer
__global__ void __launch_bounds__(RED_THREADS)
k_pack_half(const double2* __restrict__ state, double2* __restrict__ buf, uint64_t first, uint64_t n, int lbit, int want) {
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  for (uint64_t h = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; h < n; h += stride) {
    const uint64_t k = first + h;
    const uint64_t i = ((k >> lbit) << (lbit + 1)) | (k & ((1ULL << lbit) - 1ULL)) | ((uint64_t)want << lbit);
    buf[h] = state[i];
  }
}
__global__ void __launch_bounds__(RED_THREADS)
k_unpack_half(double2* __restrict__ state, const double2* __restrict__ buf, uint64_t first, uint64_t n, int lbit, int want) {
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  for (uint64_t h = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; h < n; h += stride) {
    const uint64_t k = first + h;
    const uint64_t i = ((k >> lbit) << (lbit + 1)) | (k & ((1ULL << lbit) - 1ULL)) | ((uint64_t)want << lbit);
    state[i] = buf[h];
  }
}
cudaError_t launch_pack_half(const double2* state, double2* buf, uint64_t first, uint64_t n, int lbit, int want, int grid, cudaStream_t s) {
  k_pack_half<<<grid, RED_THREADS, 0, s>>>(state, buf, first, n, lbit, want);
  return cudaGetLastError();
}
cudaError_t launch_unpack_half(double2* state, const double2* buf, uint64_t first, uint64_t n, int lbit, int want, int grid, cudaStream_t s) {
  k_unpack_half<<<grid, RED_THREADS, 0, s>>>(state, buf, first, n, lbit, want);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ in-place qubit exchange over peer-mapped memory (multi-GPU)
// Swaps k global physical bits with k local physical bits in ONE pass over NVLink, without staging buffers.  With g = the
// values of this rank's k exchanged rank bits and v = the values of an amplitude's k exchanged local bits, the amplitude
// (rank g, local v) trades places with (rank v, local g); amplitudes with v == g stay.  Of the 2^k ranks that differ only in
// the exchanged rank bits, each pair (g, v) owns a pair of sub-blocks; the pair's two ranks split the work by the top bit
// of the remaining index, so every rank reads and writes the same number of remote bytes and both NVLink directions carry
// (2^k - 1) / 2^k of a slice - against k / 2 slices for k pairwise half-slice exchanges.  One work item = one amplitude
// pair: local load + remote load, local store + remote store (16 bytes each, consecutive threads on consecutive amplitudes;
// the lowest exchanged local bit is >= 4, so a warp's 512 bytes are contiguous on both sides).  The caller brackets the launch
// with stream-ordered barriers across the ranks (nobody touches a peer's slice before its earlier kernels are done, nobody
// reads its own slice before the peers' writes have landed).
__global__ void __launch_bounds__(RED_THREADS)
k_swap_global(double2* __restrict__ mine, SwapPeers peers, SwapBits sb, uint32_t g, uint64_t n_rest_half) {
  const uint32_t k = (uint32_t)sb.k;
  const uint64_t per_partner = n_rest_half;                       // work items per partner (a power of two)
  const uint64_t total = per_partner * ((1ull << k) - 1ull);
  const uint64_t chunk = per_partner < (1ull << 13) ? per_partner : (1ull << 13);   // 128 KiB per partner visit
  const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
  // Work item -> (partner, position).  Partners are visited block-cyclically (chunks of `chunk` items) in the XOR order
  // v = g ^ s, s = 1 .. 2^k - 1, so that at any moment the CTAs of a rank talk to all partners at once and the traffic
  // into every rank comes from all of its partners at once.  (Walking the partners one after the other in the same order
  // on every rank makes all ranks hit the same receiver at the same time: 342 GB/s per direction on 8 GPUs instead of the
  // 697 GB/s of a single pair, profiles/r2e_bench_8gpu.log.)
  const uint64_t nparts = (1ull << k) - 1ull;
  auto locate = [&](uint64_t w, uint64_t& x, uint64_t& xp, double2*& peer) {
    const uint64_t c = w / chunk, in = w - c * chunk;
    const uint32_t s = 1u + (uint32_t)(c % nparts);
    uint64_t rest = (c / nparts) * chunk + in;
    const uint32_t v = g ^ s;                                     // partner's value of the exchanged bits
    // the pair's lower rank takes the items whose top rest bit is 0, the higher rank those with 1
    if (g > v) rest |= n_rest_half;
    uint64_t base = rest;
#pragma unroll
    for (int j = 0; j < MAX_SWAP_BITS; ++j) if (j < sb.k) base = ((base >> sb.lpos[j]) << (sb.lpos[j] + 1)) | (base & ((1ull << sb.lpos[j]) - 1ull));
    uint64_t dv = 0, dg = 0;
#pragma unroll
    for (int j = 0; j < MAX_SWAP_BITS; ++j)
      if (j < sb.k) { dv |= (uint64_t)((v >> sb.pair[j]) & 1u) << sb.lpos[j]; dg |= (uint64_t)((g >> sb.pair[j]) & 1u) << sb.lpos[j]; }
    x = base | dv; xp = base | dg;
    peer = peers.p[v];
  };
  uint64_t w = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  for (; w + 3 * stride < total; w += 4 * stride) {
    uint64_t x[4], xp[4]; double2* pr[4]; double2 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) locate(w + u * stride, x[u], xp[u], pr[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = __ldcs(mine + x[u]); b[u] = __ldcg(pr[u] + xp[u]); }
#pragma unroll
    for (int u = 0; u < 4; ++u) { __stcs(mine + x[u], b[u]); __stcg(pr[u] + xp[u], a[u]); }
  }
  for (; w < total; w += stride) {
    uint64_t x, xp; double2* pr;
    locate(w, x, xp, pr);
    const double2 a = __ldcs(mine + x), b = __ldcg(pr + xp);
    __stcs(mine + x, b); __stcg(pr + xp, a);
  }
}
cudaError_t launch_swap_global(double2* mine, const SwapPeers& peers, const SwapBits& sb, uint32_t g, uint64_t n_rest_half, int grid, cudaStream_t s) {
  k_swap_global<<<grid, RED_THREADS, 0, s>>>(mine, peers, sb, g, n_rest_half);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ P2: small dense complex linear algebra
// (domain/math/protocols.clj MatrixAlgebra subset; row-major interleaved)
__global__ void k_la_matmul(const double2* __restrict__ A, const double2* __restrict__ B, uint64_t m, uint64_t k, uint64_t n, double2* __restrict__ Cm) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * n) return;
  const uint64_t r = idx / n, c = idx % n;
  double2 acc{0.0, 0.0};
  for (uint64_t t = 0; t < k; ++t) { const double2 p = cmul(A[r * k + t], B[t * n + c]); acc.x += p.x; acc.y += p.y; }
  Cm[idx] = acc;
}
__global__ void k_la_kron(const double2* __restrict__ A, uint64_t ar, uint64_t ac, const double2* __restrict__ B, uint64_t br, uint64_t bc, double2* __restrict__ Cm) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t R = ar * br, Cc = ac * bc;
  if (idx >= R * Cc) return;
  const uint64_t r = idx / Cc, c = idx % Cc;
  Cm[idx] = cmul(A[(r / br) * ac + (c / bc)], B[(r % br) * bc + (c % bc)]);
}
// out[i*m + j] = x_i * conj(y_j)
__global__ void k_la_outer(const double2* __restrict__ x, const double2* __restrict__ y, uint64_t n, uint64_t m, double2* __restrict__ Cm) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const double2 a = x[idx / m], b = y[idx % m];
  Cm[idx] = double2{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}
__global__ void k_la_axpby(double2 alpha, const double2* __restrict__ x, double2 beta, const double2* __restrict__ y, uint64_t n, double2* __restrict__ out) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double2 r = cmul(alpha, x[idx]);
  if (y) { const double2 t = cmul(beta, y[idx]); r.x += t.x; r.y += t.y; }
  out[idx] = r;
}
__global__ void k_la_trace(const double2* __restrict__ A, uint64_t n, double* __restrict__ partials) {
  __shared__ double sm[2 * 32];
  double v[2] = {0.0, 0.0};
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) { const double2 a = A[i * n + i]; v[0] += a.x; v[1] += a.y; }
  block_sum<2>(v, sm);
  if (threadIdx.x == 0) { partials[0] = v[0]; partials[1] = v[1]; }
}

static inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + 255) / 256); }
cudaError_t launch_la_matmul(const double2* A, const double2* B, uint64_t m, uint64_t k, uint64_t n, double2* C, cudaStream_t s) {
  k_la_matmul<<<blocks_for(m * n), 256, 0, s>>>(A, B, m, k, n, C); return cudaGetLastError();
}
cudaError_t launch_la_kron(const double2* A, uint64_t ar, uint64_t ac, const double2* B, uint64_t br, uint64_t bc, double2* C, cudaStream_t s) {
  k_la_kron<<<blocks_for(ar * br * ac * bc), 256, 0, s>>>(A, ar, ac, B, br, bc, C); return cudaGetLastError();
}
cudaError_t launch_la_outer(const double2* x, const double2* y, uint64_t n, uint64_t m, double2* C, cudaStream_t s) {
  k_la_outer<<<blocks_for(n * m), 256, 0, s>>>(x, y, n, m, C); return cudaGetLastError();
}
cudaError_t launch_la_axpby(double2 alpha, const double2* x, double2 beta, const double2* y, uint64_t n, double2* out, cudaStream_t s) {
  k_la_axpby<<<blocks_for(n), 256, 0, s>>>(alpha, x, beta, y, n, out); return cudaGetLastError();
}
cudaError_t launch_la_trace(const double2* A, uint64_t n, double* partials, cudaStream_t s) {
  k_la_trace<<<1, 256, 0, s>>>(A, n, partials); return cudaGetLastError();
}

}  // namespace qcb
