// TMA operation rate vs box size and number of issuing warps/lanes (why the tile mover is not TMA).
// Every CTA (1 per SM) repeatedly loads a 64 KB tile from HBM as 65536/box_bytes tensor copies and stores it back.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_rate(const __grid_constant__ CUtensorMap tmap, int rows_per_box, int warps, int lanes, int tiles, int do_store,
                                                 uint64_t rows_total, long long* cycles) {
  extern __shared__ unsigned char dyn[];
  unsigned char* sm = dyn + ((1024u - (s32(dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[b]))); }
  __syncthreads();
  const int boxes = 512 / rows_per_box;                    // 512 rows of 128 B = 64 KB
  const uint32_t box_bytes = rows_per_box * 128;
  long long t0 = clock64();
  if (warp < warps && lane < lanes) {
    const int me = warp * lanes + lane, nthr = warps * lanes;
    for (int t = 0; t < tiles; ++t) {
      const int b = t & 1;
      const uint64_t row0 = ((uint64_t)(blockIdx.x + (uint64_t)t * gridDim.x) * 512) % rows_total;
      if (me == 0) asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(s32(&bar[b])), "r"(65536u) : "memory");
      for (int x = me; x < boxes; x += nthr)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sm) + b * 65536u + x * box_bytes), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"((int)(row0 + (uint64_t)x * rows_per_box)), "r"(s32(&bar[b])) : "memory");
      // wait for the tile, then store it back
      uint32_t ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar[b])), "r"((t >> 1) & 1) : "memory");
      if (do_store) {
        for (int x = me; x < boxes; x += nthr)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"((int)(row0 + (uint64_t)x * rows_per_box)), "r"(s32(sm) + b * 65536u + x * box_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const int n = 28;
  const uint64_t amps = 1ull << n, rows = amps / 8;
  double* state; cudaMalloc(&state, amps * 16); cudaMemset(state, 0, amps * 16);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, tiles = 256;
  long long* cyc; cudaMalloc(&cyc, sms * 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024);
  for (int R : {2, 4, 8, 32, 128, 256}) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {16, rows}; cuuint64_t gstr[1] = {128}; cuuint32_t box[2] = {16, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, state, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int st : {0, 1}) for (auto wl : {std::pair<int,int>{1, 1}, {1, 32}, {4, 1}, {4, 32}}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_rate<<<sms, 128, 2 * 65536 + 1024>>>(tm, R, wl.first, wl.second, 8, st, rows, cyc);
      cudaEventRecord(e0);
      k_rate<<<sms, 128, 2 * 65536 + 1024>>>(tm, R, wl.first, wl.second, tiles, st, rows, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t e = cudaGetLastError();
      double bytes = (double)sms * tiles * 65536.0 * (st ? 2 : 1);
      double ops = (double)tiles * (512 / R) * (st ? 2 : 1);
      printf("box %5d B  %s  warps %d lanes %2d : %7.3f ms  %7.1f GB/s  %6.1f cycles/op/SM  %s\n", R * 128, st ? "load+store" : "load only ", wl.first, wl.second, ms,
             bytes / ms / 1e6, ms * 1e-3 * 1.965e9 / ops, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
