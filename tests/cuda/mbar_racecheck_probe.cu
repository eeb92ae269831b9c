// mbar_racecheck_probe.cu — does compute-sanitizer's racecheck model mbarrier ordering?  Warp 0 writes a shared array and
// arrives on an mbarrier (release); warp 1 waits on the barrier's phase (acquire) and reads the array.  This is correctly
// synchronised under the PTX memory model; if racecheck still reports a hazard here, its reports on k_tile_stage's
// mbarrier-ordered accesses (tile ring: full[] / done[]) are the same tool limitation.  A second kernel does the same
// hand-over with bar.sync (which racecheck does model) as the control.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_mbar(int* out) {
  __shared__ int data[32];
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  __syncthreads();
  for (int it = 0; it < 4; ++it) {
    if (warp == 0) {
      data[lane] = it * 100 + lane;
      __syncwarp();
      if (lane == 0) asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(&bar)) : "memory");
    } else {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
      out[it * 32 + lane] = data[lane];
    }
    __syncthreads();      // keeps the next iteration's write behind this iteration's read (not the pair under test)
  }
}

__global__ void k_barsync(int* out) {
  __shared__ int data[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = 0; it < 4; ++it) {
    if (warp == 0) data[lane] = it * 100 + lane;
    __syncthreads();
    if (warp == 1) out[it * 32 + lane] = data[lane];
    __syncthreads();
  }
}

// Third kernel: the hand-over through a NAMED barrier that only part of the CTA takes part in (bar.sync 1, 64 inside a
// 128-thread block) - how the consumer groups of k_tile_stage separate their rounds.
__global__ void k_named(int* out) {
  __shared__ int data[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= 2) return;                       // warps 2, 3 never touch barrier 1
  for (int it = 0; it < 4; ++it) {
    if (warp == 0) data[lane] = it * 100 + lane;
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (warp == 1) out[it * 32 + lane] = data[lane];
    asm volatile("bar.sync 1, 64;" ::: "memory");
  }
}

// Fourth kernel: the CTA shape of k_tile_stage - 384 threads; warps 0-3 hand a shared array around through bar.sync 1, 128,
// warps 4-7 do the same on their own array through bar.sync 2, 128 at their own pace, warps 8-11 wait on an mbarrier that
// warp 0 completes at the end.  The hand-over under test is a rotation: in round r warp w writes its quarter, after the
// barrier it reads the quarter of warp (w + 1) % 4 - correctly synchronised by the group's named barrier alone.
__global__ void k_groups(int* out) {
  __shared__ int data[2][128];
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = warp >> 2, gw = warp & 3;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  __syncthreads();
  if (grp == 2) {                               // the "movers": parked on the mbarrier
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    return;
  }
  int acc = 0;
  for (int r = 0; r < 6 + 3 * grp; ++r) {       // the two groups run different numbers of rounds
    data[grp][gw * 32 + lane] = r * 1000 + gw * 32 + lane;
    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
    acc += data[grp][((gw + 1) & 3) * 32 + lane];
    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
  }
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(&bar)) : "memory");
}

int main() {
  int* d; cudaMalloc(&d, 384 * sizeof(int));
  k_groups<<<1, 384>>>(d); cudaDeviceSynchronize();
  k_barsync<<<1, 64>>>(d); cudaDeviceSynchronize();
  k_named<<<1, 128>>>(d); cudaDeviceSynchronize();
  k_mbar<<<1, 64>>>(d); cudaDeviceSynchronize();
  int h[128]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int it = 0; it < 4; ++it) for (int l = 0; l < 32; ++l) bad += h[it * 32 + l] != it * 100 + l;
  printf("mbar probe: %s (%d wrong)\n", bad ? "WRONG VALUES" : "values ok", bad);
  return bad != 0;
}
