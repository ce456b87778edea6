// Branch-free fp64 helpers for the RHS hot loops.
//
// The rollout kernels are bound by the FP64 pipe (profiles/), and the library exp()/division
// carry special-case branches that (a) cost issue slots and (b) split basic blocks so the two environments a
// thread carries cannot be interleaved by the scheduler.  These versions are straight-line code with
// 1-2 ulp error, far inside the 1e-6 parity budget, and they propagate NaN so a diverged environment
// still reaches the status check.
#pragma once

namespace pcg {

#ifndef PCG_DEV
#define PCG_DEV __device__ __forceinline__
#endif
#define PCG_HD __host__ __device__ __forceinline__

// 1/a: MUFU.RCP64H seed (2^-23) + two Newton steps -> ~1 ulp for normal a.
PCG_DEV double rcp_nr(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// exp(x), saturating at 2^+-1000 instead of overflowing, NaN-propagating.
// n = rint(x log2 e) by the 1.5*2^52 shift, r = x - n ln2 (two-part ln2), degree-12 Taylor polynomial
// in Estrin form (dependency depth 4 instead of Horner's 12), scaled by 2^n built in the exponent field.
// |r| <= 0.3466 -> truncation error 1.7e-16 relative.
PCG_DEV double exp_sat(double x) {
  constexpr double SHIFT = 6755399441055744.0;              // 1.5 * 2^52
  const double t = fma(x, 1.4426950408889634074, SHIFT);
  int n = __double2loint(t);
  const double nf = t - SHIFT;
  double r = fma(nf, -6.93147180369123816490e-01, x);
  r = fma(nf, -1.90821492927058770002e-10, r);
  n = max(-1000, min(1000, n));
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = 1.0 + r;                                // c0 + c1 r
  const double a1 = fma(1.0 / 6, r, 0.5);
  const double a2 = fma(1.0 / 120, r, 1.0 / 24);
  const double a3 = fma(1.0 / 5040, r, 1.0 / 720);
  const double a4 = fma(1.0 / 362880, r, 1.0 / 40320);
  const double a5 = fma(1.0 / 39916800, r, 1.0 / 3628800);
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(a3, r2, a2);
  const double b2 = fma(a5, r2, a4);
  const double d0 = fma(b1, r4, b0);
  const double d1 = fma(1.0 / 479001600, r4, b2);
  const double p = fma(d1, r8, d0);
  return p * __hiloint2double((n + 1023) << 20, 0);
}

}  // namespace pcg
