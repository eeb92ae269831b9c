// Micro-benchmarks that shape the tile kernel:
//  (1) dependent-issue latency of DMMA (m8n8k4 chain, one warp);  (2) m16n8k16 throughput vs warps per SM partition and
//  independent accumulators per warp;  (3) do DFMA (fp64 pipe) and DMMA (tensor pipe) overlap when issued by different warps?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_latency(double* out, long long* cyc, int iters) {
  double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 2e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NACC>
__global__ void k_mma(double* out, int iters) {
  double c[NACC][4], a[8], b[4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma16816(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA (each warp does the same work as in 0/1)
__global__ void k_mix(double* out, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  double s = 0;
  if (do_mma) {
    double c[2][4], a[8], b[4];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
    for (int it = 0; it < iters; ++it) { mma16816(c[0], a, b); mma16816(c[1], a, b); }     // 2 * 2048 MAC per warp-iteration
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  } else {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    const double a = threadIdx.x * 1e-3, b = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)                                                           // 8*16*32 = 4096 MAC per warp-iteration
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
    }
    for (int i = 0; i < 16; ++i) s += c[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
  long long* cyc; cudaMalloc(&cyc, 8);
  { const int iters = 4096; k_latency<<<1, 32>>>(out, cyc, iters); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DMMA m8n8k4: %.1f cycles per instruction\n", (double)h / iters); }
  const int iters = 2048;
  for (int wps : {4, 8, 16, 32}) {           // warps per SM (1 CTA per SM)
    auto report = [&](const char* name, int nacc, float ms) {
      double flops = (double)sms * wps * iters * nacc * 2.0 * 2048;
      printf("m16n8k16 warps/SM=%2d acc/warp=%d (%s): %.2f TFLOP/s\n", wps, nacc, name, flops / (ms * 1e-3) / 1e12);
    };
    report("1", 1, time_it([&] { k_mma<1><<<sms, wps * 32>>>(out, iters); }));
    report("2", 2, time_it([&] { k_mma<2><<<sms, wps * 32>>>(out, iters); }));
    report("4", 4, time_it([&] { k_mma<4><<<sms, wps * 32>>>(out, iters); }));
  }
  for (int wps : {8, 16, 32}) {
    float t0 = time_it([&] { k_mix<<<sms, wps * 32>>>(out, iters, 0); });
    float t1 = time_it([&] { k_mix<<<sms, wps * 32>>>(out, iters, 1); });
    float t2 = time_it([&] { k_mix<<<sms, wps * 32>>>(out, iters, 2); });
    double f = (double)sms * wps * iters * 2.0 * 4096;
    printf("mix warps/SM=%2d: all-DMMA %.3f ms (%.1f TF)  all-DFMA %.3f ms (%.1f TF)  half/half %.3f ms (%.1f TF)\n", wps, t0, f / t0 / 1e9, t1,
           f / t1 / 1e9, t2, f / t2 / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
