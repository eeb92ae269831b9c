// Micro-benchmark: FP64 tensor-core (mma.sync DMMA) vs DFMA throughput on this GPU.
// Decides whether the fused dense block of k_tile_stage should use fp64 tensor cores.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma884(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = threadIdx.x * 1e-3, b = 1.0000001;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 4096;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double warps = (double)blocks * threads / 32;
  float ms = time_it([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
  printf("dmma m8n8k4   : %.2f TFLOP/s\n", warps * iters * 8 * (2.0 * 8 * 8 * 4) / (ms * 1e-3) / 1e12);
  ms = time_it([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
  printf("dmma m16n8k16 : %.2f TFLOP/s\n", warps * iters * 4 * (2.0 * 16 * 8 * 16) / (ms * 1e-3) / 1e12);
  ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  printf("dfma          : %.2f TFLOP/s\n", warps * 32 * iters * 16 * 2.0 / (ms * 1e-3) / 1e12);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
