// Probe of the TMA 128-byte swizzle semantics the tile kernel relies on: for a box of R rows x 128 bytes
// copied to shared memory at byte offset D (multiple of 128) from a 1024-byte aligned base, where does 16-byte chunk k of
// row r land?  Expected: at row D/128 + r, chunk k ^ ((D/128 + r) & 7)  (the XOR uses the shared-memory ADDRESS bits 7..9).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int row0, uint32_t dst_off, uint32_t box_bytes, double* out, uint32_t out_doubles) {
  extern __shared__ unsigned char dyn[];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(dyn);
  uint32_t pad = (1024u - (base & 1023u)) & 1023u;
  unsigned char* sm = dyn + pad;
  __shared__ __align__(8) uint64_t bar;
  uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
  double* smd = reinterpret_cast<double*>(sm);
  for (uint32_t i = threadIdx.x; i < out_doubles; i += blockDim.x) smd[i] = -1.0;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar_s), "r"(box_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(sm) + dst_off), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(row0), "r"(bar_s) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar_s) : "memory");
  for (uint32_t i = threadIdx.x; i < out_doubles; i += blockDim.x) out[i] = smd[i];
}

int main() {
  const int n = 16;                       // 2^16 amplitudes
  const size_t amps = 1u << n;
  double* state; cudaMalloc(&state, amps * 16);
  double* h = (double*)malloc(amps * 16);
  for (size_t i = 0; i < amps * 2; ++i) h[i] = (double)i;    // double index = 2*amp + comp
  cudaMemcpy(state, h, amps * 16, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const uint32_t out_doubles = 4096 / 8 * 4;   // dump 16 KB? keep 2048 doubles = 16 KB
  double* out; cudaMalloc(&out, out_doubles * 8);
  double* ho = (double*)malloc(out_doubles * 8);
  for (int R : {1, 2, 8}) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {16, amps / 8}; cuuint64_t gstr[1] = {128}; cuuint32_t box[2] = {16, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, state, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("R=%d encode rc=%d\n", R, (int)r);
    for (uint32_t D : {0u, 128u, 256u, 384u, 512u, 1024u, 1280u}) {
      const int row0 = 40;
      k_probe<<<1, 128, 16384 + 2048>>>(tm, row0, D, R * 128, out, out_doubles);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  R=%d D=%u: %s\n", R, D, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho, out, out_doubles * 8, cudaMemcpyDeviceToHost);
      int ok_abs = 1, ok_rel = 1, ok_lin = 1, found = 0;
      for (int rr = 0; rr < R; ++rr) for (int k = 0; k < 8; ++k) {
        double want = (double)(((size_t)(row0 + rr) * 8 + k) * 2);          // re of amplitude (row, chunk k)
        int srow = D / 128 + rr;
        auto at = [&](int chunk) { return ho[(size_t)srow * 16 + chunk * 2]; };
        if (at(k ^ (srow & 7)) != want) ok_abs = 0;
        if (at(k ^ (rr & 7)) != want) ok_rel = 0;
        if (at(k) != want) ok_lin = 0;
        for (uint32_t i = 0; i < out_doubles; ++i) if (ho[i] == want) { ++found; break; }
      }
      printf("  R=%d D=%4u: found %d/%d  swizzle by smem address: %s  by row-in-box: %s  linear: %s\n", R, D, found, R * 8, ok_abs ? "YES" : "no",
             ok_rel ? "YES" : "no", ok_lin ? "YES" : "no");
    }
  }
  return 0;
}
