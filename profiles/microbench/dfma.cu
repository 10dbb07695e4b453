// Microbenchmark: FP64 FMA throughput and dependent-issue latency on B200 (used to size ILP / occupancy of the fused kernels)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, double a, double b, int iters)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, double *d)
{
  int iters = 4096;
  int threads = 32 * warps_per_sm, blocks = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<blocks, threads>>>(d, 1.0000001, 1e-9, 16);
  cudaEventRecord(e0);
  k<ILP><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = (double)blocks * threads * iters * ILP;
  double cycles = ms * 1e-3 * 1.965e9;
  printf("ILP %d warps/SM %2d : %.3f ms  %.1f TFLOP/s  FMA/clk/SM %.1f  cycles per dependent step per warp %.1f\n", ILP, warps_per_sm, ms,
         2 * fmas / ms / 1e9, fmas / cycles / 148, cycles / iters);
}
int main()
{
  double *d; cudaMalloc(&d, 148 * 1024 * 8);
  for (int w : {1, 2, 4, 8, 12, 16, 32}) run<1>(w, d);
  for (int w : {4, 12, 16}) run<2>(w, d);
  for (int w : {4, 12, 16}) run<4>(w, d);
  for (int w : {4, 12, 16, 32}) run<8>(w, d);
  return 0;
}
