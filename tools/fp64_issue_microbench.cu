// Does an FP64 warp instruction occupy the scheduler's issue port for its whole ~2.2-cycle dispatch, or only the FP64
// pipe?  One CTA on one SM; warps of role D run 4 independent DFMA chains, warps of role I run 4 independent integer
// chains (LOP3/IADD3 mix that ptxas cannot fold).  Rates are reported per SM sub-partition and cycle:
//   D alone, I alone, D and I together (same warp count each), and one warp type interleaving both 1:1.
// If DFMA only held its pipe, "together" would show D at its solo rate and I filling the remaining issue slots
// (d + i ~ 1); if it holds the issue port, d * 2.2 + i ~ 1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_issue_microbench tools/fp64_issue_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mix(double* outd, unsigned* outi, long long* cyc, int n_d, int n_i, int interleave, int iters, double a, double b, unsigned m) {
  const int warp = threadIdx.x >> 5;
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  unsigned y0 = threadIdx.x, y1 = y0 + 7, y2 = y0 + 13, y3 = y0 + 29;
  __syncthreads();
  const long long t0 = clock64();
  if (interleave) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        x0 = fma(x0, a, b); y0 = (y0 ^ m) + y1;
        x1 = fma(x1, a, b); y1 = (y1 ^ m) + y2;
        x2 = fma(x2, a, b); y2 = (y2 ^ m) + y3;
        x3 = fma(x3, a, b); y3 = (y3 ^ m) + y0;
      }
    }
  } else if (warp < n_d) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
    }
  } else if (warp < n_d + n_i) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) { y0 = (y0 ^ m) + y1; y1 = (y1 ^ m) + y2; y2 = (y2 ^ m) + y3; y3 = (y3 ^ m) + y0; }
    }
  }
  const long long t1 = clock64();
  outd[threadIdx.x] = x0 + x1 + x2 + x3;
  outi[threadIdx.x] = y0 + y1 + y2 + y3;
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

int main() {
  double* od; unsigned* oi; long long* cyc;
  cudaMalloc(&od, 1024 * 8); cudaMalloc(&oi, 1024 * 4); cudaMalloc(&cyc, 32 * 8);
  const int iters = 4000;
  auto run = [&](const char* name, int n_d, int n_i, int inter) {
    const int warps = inter ? n_d : n_d + n_i;
    mix<<<1, warps * 32>>>(od, oi, cyc, n_d, n_i, inter, iters, 0.999999, 1e-7, 0x5bd1e995u);
    long long h[32];
    cudaMemcpy(h, cyc, warps * 8, cudaMemcpyDeviceToHost);
    double cd = 0, ci = 0;
    for (int w = 0; w < warps; ++w) { if (inter || w < n_d) cd = h[w] > cd ? h[w] : cd; if (inter || w >= n_d) ci = h[w] > ci ? h[w] : ci; }
    const double per_w = 32.0 * iters;                     // DFMA (or int pairs) per warp
    const double d = (inter ? warps : n_d) * per_w / (cd > 0 ? cd : 1) / 4, i = (inter ? warps : n_i) * per_w * 2 / (ci > 0 ? ci : 1) / 4;
    printf("%-34s DFMA %.3f /cycle/SMSP   INT %.3f /cycle/SMSP   (d*2.19 + i = %.3f)\n", name, n_d || inter ? d : 0.0, n_i || inter ? i : 0.0,
           (n_d || inter ? d : 0.0) * 2.19 + (n_i || inter ? i : 0.0));
  };
  run("16 warps DFMA only", 16, 0, 0);
  run("16 warps INT only", 0, 16, 0);
  run("8 warps DFMA only", 8, 0, 0);
  run("8 warps INT only", 0, 8, 0);
  run("8 DFMA + 8 INT warps", 8, 8, 0);
  run("16 DFMA + 16 INT warps", 16, 16, 0);
  run("16 warps, DFMA:INT pairs 1:1", 16, 0, 1);
  run("32 warps, DFMA:INT pairs 1:1", 32, 0, 1);
  return 0;
}
