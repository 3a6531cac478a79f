// Microbenchmark: throughput of warp-aggregated atomics on ONE address (the ray-queue counter).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ret(unsigned* ctr, unsigned* out, int iters) {
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    unsigned b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(ctr, 32u);
    b = __shfl_sync(0xffffffffu, b, 0);
    acc += b;
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}
__global__ void k_red(unsigned long long* ctr, int iters) {
  for (int i = 0; i < iters; ++i)
    if ((threadIdx.x & 31) == 0) atomicAdd(ctr, 1ull);
}
int main() {
  unsigned *c, *o; unsigned long long* c64;
  cudaMalloc(&c, 1024); cudaMalloc(&o, 4); cudaMalloc(&c64, 8);
  cudaMemset(c, 0, 1024); cudaMemset(c64, 0, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks : {148, 148 * 4, 148 * 8}) {
    for (int mode = 0; mode < 2; ++mode) {
      const int iters = 200, threads = 128;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_ret<<<blocks, threads>>>(c, o, iters); else k_red<<<blocks, threads>>>(c64, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double n = (double)blocks * (threads / 32) * iters;
      printf("%s blocks=%d: %.0f atomics in %.3f ms -> %.2f ns/atomic (%.1f M/s)\n", mode ? "RED u64 " : "ATOM ret", blocks, n, ms, ms * 1e6 / n, n / ms / 1e3);
    }
  }
  return 0;
}
