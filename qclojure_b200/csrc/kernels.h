// kernels.h — launch wrappers of kernels.cu (host-callable), shared constants and small POD params.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qcb {

constexpr int TILE_THREADS = 256;
constexpr int RED_THREADS = 256;
constexpr int SAMPLE_THREADS = 256;
constexpr int SAMPLE_PER_THREAD = 16;
constexpr int SAMPLE_CHUNK = SAMPLE_THREADS * SAMPLE_PER_THREAD;   // amplitudes per sampling chunk
constexpr int EXPECT_TERMS = 16;                                   // Pauli terms per expectation pass
constexpr int MAX_MEASURE_BITS = 24;                               // qubits per :measure op (two histogram passes of <= 12 bits)
constexpr int MAX_HIST_BITS = 12;                                  // bins of one marginal pass: 2^12 doubles per warp in shared memory

struct ExpectTerms { int n; uint64_t zmask[EXPECT_TERMS]; double pr[EXPECT_TERMS], pi[EXPECT_TERMS]; };
struct Mat2 { double m[8]; };
struct BitList { int n; int pos[MAX_MEASURE_BITS]; };
struct GroverMarks { int n; uint64_t idx[8]; };                    // marked basis states of the phase oracles (local indices)
// multi-qubit exchange (k_swap_global): lpos[j] = j-th exchanged LOCAL bit position, ASCENDING; pair[j] = which bit of the
// group value (the exchanged rank bits, in ascending rank-bit order) it trades places with; peers.p[v] = slice of the rank whose
// exchanged rank bits have the value v (p[own value] unused)
constexpr int MAX_SWAP_BITS = 3;
struct SwapBits { int k; int lpos[MAX_SWAP_BITS]; int pair[MAX_SWAP_BITS]; };
struct SwapPeers { double2* p[1 << MAX_SWAP_BITS]; };

// TMA tensor maps of one state allocation, indexed by run bits c (box = 2^(c-3) rows of 128 bytes); opaque 128-byte
// CUtensorMap objects so that this header does not need <cuda.h>.
struct TileMaps {
  alignas(64) unsigned char map[12][128];
  bool valid[12];
};
// Builds the maps for a state of 2^n_local amplitudes at `state` (driver entry point resolved at run time).
cudaError_t build_tile_maps(double2* state, int n_local, TileMaps* out);
cudaError_t launch_tile_stage(double2* state, const uint64_t* stage_dev, const uint64_t* stage_host, uint32_t stage_words,
                              const double* dev_vals, int num_sms, cudaStream_t stream, uint64_t* out_active, const TileMaps* maps);
void tile_prof_dump();   // prints the cycle accounting of k_tile_stage (only in a PROFILE=1 build)
void tile_trace_dump();  // prints the timeline trace of k_tile_stage (only in a -DQCB_TILE_TRACE build)
cudaError_t launch_set_amp(double2* state, uint64_t idx, double re, double im, cudaStream_t s);
cudaError_t launch_reduce(const double2* state, uint64_t count, int mode, double* partials, int grid, cudaStream_t s);
cudaError_t launch_finalize(const double* partials, uint32_t nparts, uint32_t K, int post, double param, double* out, cudaStream_t s);
cudaError_t launch_grover_step(double2* state, uint64_t count, const double* coef, const GroverMarks& marks, double* partials,
                               int grid, cudaStream_t s);
cudaError_t launch_scale_dev(double2* state, uint64_t count, const double* coef, int grid, cudaStream_t s);
cudaError_t launch_probabilities(const double2* state, uint64_t offset, uint64_t count, double* out, int grid, cudaStream_t s);
cudaError_t launch_gather(const double2* state, const uint64_t* idx, uint64_t n, double2* out, cudaStream_t s);
cudaError_t launch_inner(const double2* phi, const double2* psi, uint64_t count, double* partials, int grid, cudaStream_t s);
cudaError_t launch_expect_group(const double2* state, uint64_t count, uint64_t xmask, int pivot, uint64_t ext_or,
                                const ExpectTerms& terms, double* partials, int grid, cudaStream_t s);
cudaError_t launch_expect_1q(const double2* state, uint64_t count, int bit, const Mat2& O, double* partials, int grid, cudaStream_t s);
cudaError_t launch_marginal(const double2* state, uint64_t count, uint64_t ext_or, const BitList& bl, const BitList& filter, uint32_t filter_val,
                            double* partials, int grid, cudaStream_t s);
cudaError_t launch_collapse(double2* state, uint64_t count, uint64_t ext_or, const BitList& bl, uint32_t sel, double factor, int grid, cudaStream_t s);
cudaError_t launch_chunk_sums(const double2* state, uint64_t count, double* sums, int grid, cudaStream_t s);
cudaError_t launch_scan_inclusive(double* v, uint64_t n, cudaStream_t s);
cudaError_t launch_sample(const double2* state, uint64_t count, const double* cum_chunks, uint64_t n_chunks, const double* uniforms,
                          uint64_t n_shots, double total, double rank_offset, int first, int last, uint64_t index_offset,
                          unsigned long long* outcomes, int grid, cudaStream_t s);
cudaError_t launch_pack_half(const double2* state, double2* buf, uint64_t first, uint64_t n, int lbit, int want, int grid, cudaStream_t s);
cudaError_t launch_unpack_half(double2* state, const double2* buf, uint64_t first, uint64_t n, int lbit, int want, int grid, cudaStream_t s);
cudaError_t launch_swap_global(double2* mine, const SwapPeers& peers, const SwapBits& sb, uint32_t g, uint64_t n_rest_half, int grid, cudaStream_t s);
cudaError_t launch_la_matmul(const double2* A, const double2* B, uint64_t m, uint64_t k, uint64_t n, double2* C, cudaStream_t s);
cudaError_t launch_la_kron(const double2* A, uint64_t ar, uint64_t ac, const double2* B, uint64_t br, uint64_t bc, double2* C, cudaStream_t s);
cudaError_t launch_la_outer(const double2* x, const double2* y, uint64_t n, uint64_t m, double2* C, cudaStream_t s);
cudaError_t launch_la_axpby(double2 alpha, const double2* x, double2 beta, const double2* y, uint64_t n, double2* out, cudaStream_t s);
cudaError_t launch_la_trace(const double2* A, uint64_t n, double* partials, cudaStream_t s);

}  // namespace qcb
