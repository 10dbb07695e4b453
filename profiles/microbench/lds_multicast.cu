// Microbenchmark: cost of a warp-wide 64-bit shared-memory load on B200 as a function of the number of DISTINCT words the
// 32 lanes read (multicast).  Sizes the round-2 kernel formulation (one thread per (x, y) column: the lanes of a row read the
// same word, 5-7 distinct words per instruction) against today's (every lane a different word: two wavefronts).
// Per pattern: loads per clock and SM with 8 warps per SM, 8 independent loads in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>

// lane -> word index: `distinct` groups of consecutive lanes share a word; words are `stride` doubles apart
__global__ void k(double *out, int distinct, int stride, int iters)
{
  __shared__ double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int group = lane * distinct / 32;      // 0 .. distinct-1
  int idx = (group * stride) & 4095;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++) acc[u] += s[(idx + u * 64) & 4095];
    idx = (idx + 1) & 4095; // keeps the compiler from hoisting the loads; the pattern (distinct words, spacing) is unchanged
  }
  double t = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) t += acc[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main()
{
  double *d;
  cudaMalloc(&d, 148 * 256 * sizeof(double));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000, blocks = 148, threads = 256;
  printf("B200 LDS.64 multicast: 148 CTAs x 8 warps, 8 loads in flight per thread; SM clock attribute %d kHz\n", clk_khz);
  printf("%10s %8s %12s %22s\n", "distinct", "stride", "ms", "warp loads / clk / SM");
  const int pats[][2] = {{32, 1}, {32, 5}, {16, 1}, {8, 1}, {7, 5}, {7, 25}, {5, 1}, {5, 25}, {2, 1}, {1, 1}, {7, 16}, {8, 2}};
  for (auto &p : pats)
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, threads>>>(d, p[0], p[1], 100);
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d, p[0], p[1], iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_loads = (double)iters * 8 * (threads / 32); // per SM
    const double clks = ms * 1e-3 * 1.965e9;                       // SM clock under load on this pool (MEASURED_PEAKS.json)
    printf("%10d %8d %12.3f %22.3f\n", p[0], p[1], ms, warp_loads / clks);
  }
  return 0;
}
