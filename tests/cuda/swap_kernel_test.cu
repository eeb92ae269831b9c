// swap_kernel_test.cu — single-GPU check of k_swap_global (the multi-qubit exchange kernel) for k = 1, 2, 3 exchanged
// bits: the 2^p "ranks" of a world are p-bit-indexed slices allocated on ONE device, peer pointers are plain pointers, and
// every rank's launch runs on the same stream.  A pair of ranks splits each sub-block pair between them, so running the
// ranks' kernels one after the other gives exactly what concurrent execution over NVLink gives.  Checked against the
// definition: swapping global bit G_j with local bit l_j moves amplitude (rank, x) to (rank', x') with rank' = rank with
// bit (G_j - nl) := x's bit l_j and x' = x with bit l_j := rank's bit (G_j - nl).
// Build: nvcc -arch=sm_100a -I../../qclojure_b200/csrc swap_kernel_test.cu ../../qclojure_b200/csrc/kernels.cu -o swap_kernel_test
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "kernels.h"

using namespace qcb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main() {
  std::mt19937_64 rng(7);
  int n_cases = 0;
  for (int p = 1; p <= 3; ++p) {                       // world = 2^p ranks
    const int world = 1 << p;
    for (int nl : {10, 14, 17}) {
      const uint64_t lc = 1ull << nl;
      std::vector<double2*> d(world);
      for (int r = 0; r < world; ++r) CK(cudaMalloc(&d[r], lc * sizeof(double2)));
      for (int trial = 0; trial < 6; ++trial) {
        const int k = 1 + (int)(rng() % p);             // exchanged pairs
        // choose k distinct global bits and k distinct local bits
        std::vector<int> gb(p), lb(nl);
        for (int i = 0; i < p; ++i) gb[i] = nl + i;
        for (int i = 0; i < nl; ++i) lb[i] = i;
        std::shuffle(gb.begin(), gb.end(), rng);
        std::shuffle(lb.begin(), lb.end(), rng);
        std::vector<std::pair<int, int>> pairs(k);
        for (int j = 0; j < k; ++j) pairs[j] = {gb[j], lb[j]};
        // fill: value encodes (rank, index)
        std::vector<std::vector<double2>> h(world, std::vector<double2>(lc));
        for (int r = 0; r < world; ++r) {
          for (uint64_t x = 0; x < lc; ++x) h[r][x] = double2{(double)r, (double)x};
          CK(cudaMemcpy(d[r], h[r].data(), lc * sizeof(double2), cudaMemcpyHostToDevice));
        }
        // the same argument construction as sim.cu: do_exchange_swap
        std::vector<int> gs(k), order(k);
        for (int j = 0; j < k; ++j) gs[j] = pairs[j].first;
        std::sort(gs.begin(), gs.end());
        for (int j = 0; j < k; ++j) order[j] = j;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return pairs[a].second < pairs[b].second; });
        SwapBits sb; sb.k = k;
        for (int j = 0; j < MAX_SWAP_BITS; ++j) { sb.lpos[j] = 0; sb.pair[j] = 0; }
        for (int j = 0; j < k; ++j) {
          sb.lpos[j] = pairs[order[j]].second;
          sb.pair[j] = (int)(std::find(gs.begin(), gs.end(), pairs[order[j]].first) - gs.begin());
        }
        for (int rank = 0; rank < world; ++rank) {
          uint32_t g = 0;
          for (int i = 0; i < k; ++i) g |= (uint32_t)((rank >> (gs[i] - nl)) & 1) << i;
          SwapPeers peers;
          for (int v = 0; v < (1 << MAX_SWAP_BITS); ++v) peers.p[v] = nullptr;
          for (uint32_t v = 0; v < (1u << k); ++v) {
            if (v == g) continue;
            int r = rank;
            for (int i = 0; i < k; ++i) r = (r & ~(1 << (gs[i] - nl))) | (int)((v >> i) & 1u) << (gs[i] - nl);
            peers.p[v] = d[r];
          }
          const uint64_t n_rest_half = lc >> (k + 1);
          CK(launch_swap_global(d[rank], peers, sb, g, n_rest_half, 64, 0));
        }
        CK(cudaDeviceSynchronize());
        // verify
        uint64_t bad = 0;
        for (int r = 0; r < world; ++r) {
          std::vector<double2> out(lc);
          CK(cudaMemcpy(out.data(), d[r], lc * sizeof(double2), cudaMemcpyDeviceToHost));
          for (uint64_t x = 0; x < lc; ++x) {
            // (r, x) after the exchange holds the amplitude that was at (r0, x0): the swap is an involution
            int r0 = r; uint64_t x0 = x;
            for (auto& pr : pairs) {
              const int gbit = pr.first - nl, lbit = pr.second;
              const int rb = (r >> gbit) & 1, xb = (int)((x >> lbit) & 1);
              r0 = (r0 & ~(1 << gbit)) | (xb << gbit);
              x0 = (x0 & ~(1ull << lbit)) | ((uint64_t)rb << lbit);
            }
            if (out[x].x != (double)r0 || out[x].y != (double)x0) ++bad;
          }
        }
        ++n_cases;
        if (bad) { printf("FAIL world=%d nl=%d k=%d: %llu wrong amplitudes\n", world, nl, k, (unsigned long long)bad); return 1; }
      }
      for (int r = 0; r < world; ++r) cudaFree(d[r]);
    }
  }
  printf("swap kernel ok: %d cases\n", n_cases);
  return 0;
}
