// Micro-benchmark behind DESIGN.md section 5 (round 2): what does the instruction mix of a three-product round cost on the fp64
// pipe of one SM partition?  Registers only - no shared or global memory - W warps per SM partition, results of every product
// fed back as operands of the next iteration so that ptxas (which treats mma as a pure instruction: it deletes dead ones and
// hoists loop-invariant ones) keeps them.  The DMMA / DADD counts per loop iteration printed below are what
// `cuobjdump -sass scripts/dmma_mix` shows in each loop (re-check after changing anything: a first version of this file
// "measured" 53 TFLOP/s because a third of its DMMAs had been optimised away).
//   skew2   the skewed two-batch pipeline of k3_pp (kernels.cu): 6 DMMA + 4 DADD per batch, DADD results feed the DMMAs
//   skew0   the same without the additions
//   four    four-product form: 8 DMMA per batch, no additions
//   lock1   one-sum form (s = Br + Bi; K = Mr s; Re = -(Mr + Mi) Bi + K; Im = (Mi - Mr) Br + K), two batches in lockstep
//   lock2   the same for a paired round (second block fed from the first block's results)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_mix dmma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

struct Set { double K0, K1, s0, s1, d0, d1, r0, i0, r1, i1; };

template <int MODE>   // 0 = skew0, 1 = skew2, 4 = four
__device__ __forceinline__ void batch(Set& c, Set& n, double& pr0, double& pr1, double& pi0, double& pi1, const double (&A)[8]) {
  mma(pr0, pr1, A[2], c.s0, c.K0, c.K1);
  mma(n.K0, n.K1, A[0], n.r0, 0.0, 0.0);
  mma(pi0, pi1, A[4], c.d0, c.K0, c.K1);
  mma(n.K0, n.K1, A[1], n.r1, n.K0, n.K1);
  if (MODE == 1) { n.s0 = n.r0 + n.i0; n.d0 = n.i0 - n.r0; }
  mma(pr0, pr1, A[3], c.s1, pr0, pr1);
  if (MODE == 1) { n.s1 = n.r1 + n.i1; n.d1 = n.i1 - n.r1; }
  mma(pi0, pi1, A[5], c.d1, pi0, pi1);
  if (MODE == 4) {
    mma(pr0, pr1, A[6], c.r0, pr0, pr1);
    mma(pi0, pi1, A[7], c.i0, pi0, pi1);
  }
  c.r0 = pr0; c.i0 = pi0; c.r1 = pr1; c.i1 = pi1;
  if (MODE != 1) { c.s0 = c.r0; c.d0 = c.i0; c.s1 = c.r1; c.d1 = c.i1; }
}

template <int MODE>
__global__ void k_skew(double* out, long long* cyc, int iters) {
  double A[8];
  for (int i = 0; i < 8; ++i) A[i] = ((threadIdx.x + i) % 8 == 0) ? 0.35 : 0.0;
  Set a, b;
  a.K0 = a.K1 = 0; a.r0 = threadIdx.x * 1e-4; a.i0 = 1e-3; a.r1 = 2e-3; a.i1 = -1e-3; a.s0 = a.r0 + a.i0; a.d0 = a.i0 - a.r0; a.s1 = a.r1 + a.i1; a.d1 = a.i1 - a.r1;
  b = a; b.r0 += 1e-5;
  double pr0 = 0, pr1 = 0, pi0 = 0, pi1 = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    batch<MODE>(a, b, pr0, pr1, pi0, pi1, A);
    batch<MODE>(b, a, pr0, pr1, pi0, pi1, A);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = pr0 + pr1 + pi0 + pi1 + a.K0 + b.K0;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

struct Raw { double r0, i0, r1, i1; };
template <int PAIR>
__global__ void k_lock(double* out, long long* cyc, int iters) {
  double A[12];
  for (int i = 0; i < 12; ++i) A[i] = ((threadIdx.x + i) % 8 == 0) ? 0.35 : 0.0;
  Raw a, b;
  a.r0 = threadIdx.x * 1e-4; a.i0 = 1e-3; a.r1 = 2e-3; a.i1 = -1e-3; b = a; b.r0 += 1e-5;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    double sa0 = a.r0 + a.i0, sa1 = a.r1 + a.i1, sb0 = b.r0 + b.i0, sb1 = b.r1 + b.i1;
    double ka0, ka1, kb0, kb1, xa0, xa1, ya0, ya1, xb0, xb1, yb0, yb1;
    mma(ka0, ka1, A[0], sa0, 0.0, 0.0); mma(kb0, kb1, A[0], sb0, 0.0, 0.0);
    mma(ka0, ka1, A[1], sa1, ka0, ka1); mma(kb0, kb1, A[1], sb1, kb0, kb1);
    mma(xa0, xa1, A[2], a.i0, ka0, ka1); mma(xb0, xb1, A[2], b.i0, kb0, kb1);
    mma(ya0, ya1, A[4], a.r0, ka0, ka1); mma(yb0, yb1, A[4], b.r0, kb0, kb1);
    mma(xa0, xa1, A[3], a.i1, xa0, xa1); mma(xb0, xb1, A[3], b.i1, xb0, xb1);
    mma(ya0, ya1, A[5], a.r1, ya0, ya1); mma(yb0, yb1, A[5], b.r1, yb0, yb1);
    if (PAIR) {
      sa0 = xa0 + ya0; sa1 = xa1 + ya1; sb0 = xb0 + yb0; sb1 = xb1 + yb1;
      mma(ka0, ka1, A[6], sa0, 0.0, 0.0); mma(kb0, kb1, A[6], sb0, 0.0, 0.0);
      mma(ka0, ka1, A[7], sa1, ka0, ka1); mma(kb0, kb1, A[7], sb1, kb0, kb1);
      double ua0, ua1, va0, va1, ub0, ub1, vb0, vb1;
      mma(ua0, ua1, A[8], ya0, ka0, ka1); mma(ub0, ub1, A[8], yb0, kb0, kb1);
      mma(va0, va1, A[10], xa0, ka0, ka1); mma(vb0, vb1, A[10], xb0, kb0, kb1);
      mma(ua0, ua1, A[9], ya1, ua0, ua1); mma(ub0, ub1, A[9], yb1, ub0, ub1);
      mma(va0, va1, A[11], xa1, va0, va1); mma(vb0, vb1, A[11], xb1, vb0, vb1);
      a.r0 = ua0; a.r1 = ua1; a.i0 = va0; a.i1 = va1; b.r0 = ub0; b.r1 = ub1; b.i0 = vb0; b.i1 = vb1;
    } else {
      a.r0 = xa0; a.r1 = xa1; a.i0 = ya0; a.i1 = ya1; b.r0 = xb0; b.r1 = xb1; b.i0 = yb0; b.i1 = yb1;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a.r0 + a.r1 + a.i0 + a.i1 + b.r0 + b.i0 + b.r1 + b.i1;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class K>
static void run(const char* name, K kernel, int n_dmma, int n_dadd, int warps_per_sm, int sms, double* out, long long* cyc) {
  const int iters = 200000;
  kernel<<<sms, warps_per_sm * 32>>>(out, cyc, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kernel<<<sms, warps_per_sm * 32>>>(out, cyc, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_iter_smsp = (double)h / iters / (warps_per_sm / 4.0);      // pipe cycles of one SM partition per loop iteration
  const double tf = (double)n_dmma * iters * warps_per_sm * sms * 512.0 / (ms * 1e-3) / 1e12;
  printf("%-6s warps/SMSP %d  %2d DMMA + %d DADD per iteration: %6.1f cycles/iteration/SMSP = %5.2f per DMMA  | %5.1f TFLOP/s from the DMMAs "
         "(wall %.2f ms, clock64 at %.0f MHz)\n", name, warps_per_sm / 4, n_dmma, n_dadd, per_iter_smsp, per_iter_smsp / n_dmma, tf, ms,
         (double)h / (ms * 1e-3) / 1e6);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) {
    run("skew0", k_skew<0>, 10, 0, w, sms, out, cyc);     // ptxas folds the two C = 0 products of an iteration away
    run("skew2", k_skew<1>, 12, 8, w, sms, out, cyc);
    run("four", k_skew<4>, 13, 3, w, sms, out, cyc);
    run("lock1", k_lock<0>, 12, 4, w, sms, out, cyc);
    run("lock2", k_lock<1>, 24, 8, w, sms, out, cyc);
  }
  return 0;
}
