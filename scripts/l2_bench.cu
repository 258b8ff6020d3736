// L2 read-bandwidth microbenchmark (BASELINE.md asks for a measured L2 peak next to the HBM one):
// the grid streams a `bytes`-sized buffer (which fits in the 126 MB L2) with 128-bit L1-bypassing loads, many times;
// each element is read by exactly one thread per pass (no cross-CTA request merging).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bench l2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const uint4* __restrict__ p, size_t n, int reps, unsigned* sink) {
  unsigned acc = 0;
  for (int r = 0; r < reps; r++)
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
      size_t i = (k + (size_t)r * 7919 * 64) % n;    // shift every pass so CTAs do not re-read the lines they just fetched
      uint4 v = __ldcg(p + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  unsigned* sink; cudaMalloc(&sink, 4);
  for (size_t mb : {8, 16, 32, 64, 96, 256, 1024}) {
    size_t bytes = mb << 20, n = bytes / 16;
    uint4* p; cudaMalloc(&p, bytes); cudaMemset(p, 1, bytes);
    int reps = mb <= 96 ? 40 : 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rd<<<148 * 8, 512>>>(p, n, 2, sink);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int t = 0; t < 5; t++) {
      cudaEventRecord(a); rd<<<148 * 8, 512>>>(p, n, reps, sink); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%5zu MiB working set: %.1f GB/s read (%s)\n", mb, (double)bytes * reps / best / 1e6, mb <= 96 ? "L2-resident" : "HBM");
    cudaFree(p);
  }
  return 0;
}
