// Measures on the target GPU: dependent-issue latency of DFMA / DMUL / MUFU.RCP64H, and the per-SM DFMA
// throughput as a function of resident warps (the rollout kernels are bound by this pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8 * 64);
  long long h[148 * 64];
  const int iters = 2000;
  // latency: 1 warp, 1 chain
  dfma_chain<1><<<1, 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", (double)h[0] / (iters * 16));
  dfma_chain<2><<<1, 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 1 warp ILP2: %.2f cycles per DFMA\n", (double)h[0] / (iters * 16 * 2));
  dfma_chain<4><<<1, 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 1 warp ILP4: %.2f cycles per DFMA\n", (double)h[0] / (iters * 16 * 4));
  dfma_chain<8><<<1, 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 1 warp ILP8: %.2f cycles per DFMA\n", (double)h[0] / (iters * 16 * 8));
  // throughput per SM vs warps per SM (1 chain per warp)
  for (int warps : {4, 8, 12, 14, 16, 24, 32, 64}) {
    dfma_chain<1><<<1, warps * 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 SM, %2d warps x ILP1: %.3f warp-DFMA/cycle/SM\n", warps, (double)warps * iters * 16 / h[0]);
  }
  for (int warps : {4, 8, 12, 14, 16}) {
    dfma_chain<2><<<1, warps * 32>>>(out, cyc, 0.999999, 1e-7, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 SM, %2d warps x ILP2: %.3f warp-DFMA/cycle/SM\n", warps, (double)warps * iters * 16 * 2 / h[0]);
  }
  // whole chip
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_chain<4><<<148 * 2, 512>>>(out, cyc, 0.999999, 1e-7, iters);
  cudaEventRecord(e0);
  dfma_chain<4><<<148 * 2, 512>>>(out, cyc, 0.999999, 1e-7, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 148 * 2 * 512 * 4 * 16.0 * iters;
  printf("chip DFMA: %.2f TFLOP/s fp64 (%.3f ms)\n", flops / ms * 1e-9, ms);
  return 0;
}
