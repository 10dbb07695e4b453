// Microbenchmark: FP64 tensor-core (mma.sync.m8n8k4.f64) throughput on B200, alone and mixed with DFMA in the same warp,
// to decide whether the per-line operator contractions of the fused kernels belong on DMMA tiles.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NM independent DMMA accumulators and NF independent DFMA chains per thread and iteration
template <int NM, int NF>
__global__ void k(double *out, double a, double b, int iters)
{
  double c[NM > 0 ? NM : 1][2], x[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NM; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NF; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int i = 0; i < NM; i++) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
void run(int warps_per_sm, double *d)
{
  int iters = 2048;
  int threads = 32 * warps_per_sm, blocks = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NM, NF><<<blocks, threads>>>(d, 1.0000001, 1e-9, 16);
  cudaEventRecord(e0);
  k<NM, NF><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_iters = (double)blocks * warps_per_sm * iters;
  double mma_fmas = warp_iters * NM * 256, vec_fmas = warp_iters * NF * 32;
  double cycles = ms * 1e-3 * 1.965e9;
  printf("DMMA x%d + DFMA x%d, warps/SM %2d : %.3f ms | DMMA %.1f FMA/clk/SM (%.1f TFLOP/s) | DFMA %.1f FMA/clk/SM | cycles/iter/warp %.1f\n", NM, NF,
         warps_per_sm, ms, mma_fmas / cycles / 148, 2 * mma_fmas / ms / 1e9, vec_fmas / cycles / 148, cycles / iters);
}

int main()
{
  double *d; cudaMalloc(&d, 148 * 1024 * 8);
  for (int w : {1, 4, 8, 16, 32}) run<1, 0>(w, d); // latency / throughput, one chain
  for (int w : {4, 8, 16}) run<4, 0>(w, d);
  for (int w : {4, 16}) run<8, 0>(w, d);
  for (int w : {4, 16}) run<0, 8>(w, d);           // DFMA alone
  for (int w : {4, 8, 16}) run<4, 8>(w, d);        // do the two overlap ?  (4 DMMA = 1024 FMA, 8 DFMA = 256 FMA per warp-iter)
  for (int w : {4, 8, 16}) run<2, 16>(w, d);
  for (int w : {4, 8, 16}) run<1, 32>(w, d);
  return 0;
}
