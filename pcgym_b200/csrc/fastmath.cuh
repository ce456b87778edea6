// Branch-free fp64 helpers for the RHS hot loops and the Gaussian draws.
//
// The rollout kernels are bound by the FP64 pipe and, at 65 536 envs, by dependent-issue latency
// (profiles/).  The CUDA library exp()/log()/sincospi()/division carry special-case branches
// that (a) cost issue slots and (b) split basic blocks, so the two environments a thread carries cannot
// be interleaved by the scheduler.  These versions are straight-line code with a few ulp of error -- far
// inside the 1e-6 parity budget -- and they propagate NaN so a diverged environment still reaches the
// status check.  Polynomial coefficients live in __constant__ memory: one LDC / LDCU per 64-bit constant (sm_100
// SASS has no constant-bank operands on FP64 instructions) instead of two UMOVs per 64-bit immediate.
#pragma once

namespace pcg {

#ifndef PCG_DEV
#define PCG_DEV __device__ __forceinline__
#endif
#define PCG_HD __host__ __device__ __forceinline__

struct MathConst {
  double log2e, ln2_hi, ln2_lo, shift;
  double ex[11];      // exp(r) on |r| <= ln2/2: degree-10 interpolant at Chebyshev nodes, max rel. error 4.4e-16
  double at[10];      // 2/(2k+1): ln m = s * sum at[k] s^2k,  s = (m-1)/(m+1)
  double sn[8];       // sin(pi x)/x   in x^2, |x| <= 1/4
  double cs[8];       // cos(pi x)     in x^2
};

static __constant__ MathConst kMC = {
    1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 6755399441055744.0,
    {1.0, 1.0000000000000067, 0.5000000000000006, 0.16666666666554325, 0.04166666666657431, 0.008333333385696902,
     0.0013888888932304278, 0.0001984117023243583, 2.4801504458882045e-05, 2.764019693014132e-06,
     2.762633679135001e-07},
    {2.0, 2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19},
    // Taylor coefficients of sin(pi x)/x and cos(pi x) in x^2: (-1)^k pi^(2k+1)/(2k+1)!, (-1)^k pi^(2k)/(2k)!
    {3.141592653589793, -5.167712780049969, 2.550164039877345, -0.5992645293207919, 0.08214588661112819, -0.007370430945714348, 0.00046630280576761234, -2.1915353447830204e-05},
    {1.0, -4.934802200544679, 4.058712126416768, -1.3352627688545893, 0.23533063035889312, -0.02580689139001405, 0.001929574309403922, -0.00010463810492484565},
};

// 1/a: MUFU.RCP64H seed (2^-23) + two Newton steps -> ~1 ulp for normal a.
PCG_DEV double rcp_nr(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// sqrt(a) for a > 0: MUFU.RSQ64H seed + Newton on y = 1/sqrt(a), then one correction of a*y.
PCG_DEV double sqrt_nr(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double h = 0.5 * y;
  double e = fma(-a * y, h, 0.5);          // 1/2 - a y^2 / 2
  y = fma(y, e, y);
  h = 0.5 * y;
  e = fma(-a * y, h, 0.5);
  y = fma(y, e, y);
  const double s = a * y;
  return fma(fma(-s, s, a), 0.5 * y, s);
}

// exp(x), saturating at 2^+-1000 instead of overflowing, NaN-propagating.
// n = rint(x log2 e) by the 1.5*2^52 shift, r = x - n ln2 (two-part ln2), degree-10 near-minimax polynomial
// in Estrin form (dependency depth 4 instead of Horner's 10), scaled by 2^n built in the exponent field.
PCG_DEV double exp_sat(double x) {
  const double t = fma(x, kMC.log2e, kMC.shift);
  int n = __double2loint(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
  n = max(-1000, min(1000, n));
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = fma(kMC.ex[1], r, kMC.ex[0]);
  const double a1 = fma(kMC.ex[3], r, kMC.ex[2]);
  const double a2 = fma(kMC.ex[5], r, kMC.ex[4]);
  const double a3 = fma(kMC.ex[7], r, kMC.ex[6]);
  const double a4 = fma(kMC.ex[9], r, kMC.ex[8]);
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(a3, r2, a2);
  const double b2 = fma(kMC.ex[10], r2, a4);
  const double p = fma(b2, r8, fma(b1, r4, b0));
  return p * __hiloint2double((n + 1023) << 20, 0);
}

// Same function with the polynomial in Horner form: 10 dependent DFMA instead of 13 DFMA/DMUL at depth 4.
// Three FP64 issue slots cheaper per call; used where other warps hide the longer dependency chain.
PCG_DEV double exp_sat_horner(double x) {
  const double t = fma(x, kMC.log2e, kMC.shift);
  int n = __double2loint(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
  n = max(-1000, min(1000, n));
  double p = kMC.ex[10];
#pragma unroll
  for (int i = 9; i >= 0; --i) p = fma(p, r, kMC.ex[i]);
  return p * __hiloint2double((n + 1023) << 20, 0);
}

// Horner form with the 2^n scaling done by an integer add into the exponent field of the polynomial value (one
// integer instruction instead of a move and a DMUL).  p is in [0.70, 1.42] and |n| <= 1000, so the sum stays a
// normal number.  NOT NaN-preserving on its own: a NaN argument can come back finite.  Only for right-hand sides
// whose NaN state re-enters the derivative through another term (cstr: T and ca both appear linearly as well).
PCG_DEV double exp_sat_horner_iscale(double x) {
  const double t = fma(x, kMC.log2e, kMC.shift);
  int n = __double2loint(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
  n = max(-1000, min(1000, n));
  double p = kMC.ex[10];
#pragma unroll
  for (int i = 9; i >= 0; --i) p = fma(p, r, kMC.ex[i]);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// ln(u) for a positive NORMAL double (the Box-Muller argument is (k + 1/2) 2^-53, never 0 or denormal).
// u = m 2^e with m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716;
// odd series to s^19 (remainder < 2e-16), Estrin form.
PCG_DEV double log_pos(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;              // m in [1, 2)
  const bool big = hi >= 0x3ff6a09f;                // m >= sqrt(2) (to 20 bits): halve it
  hi = big ? hi - 0x00100000 : hi;
  e = big ? e + 1 : e;
  const double m = __hiloint2double(hi, lo);
  const double s = (m - 1.0) * rcp_nr(m + 1.0);
  const double z = s * s, z2 = z * z, z4 = z2 * z2;
  const double a0 = fma(kMC.at[1], z, kMC.at[0]);
  const double a1 = fma(kMC.at[3], z, kMC.at[2]);
  const double a2 = fma(kMC.at[5], z, kMC.at[4]);
  const double a3 = fma(kMC.at[7], z, kMC.at[6]);
  const double a4 = fma(kMC.at[9], z, kMC.at[8]);
  const double b0 = fma(a1, z2, a0);
  const double b1 = fma(a3, z2, a2);
  const double p = fma(fma(a4, z4, b1), z4, b0);
  const double ef = static_cast<double>(e);
  return fma(ef, -kMC.ln2_hi, fma(s, p, ef * -kMC.ln2_lo));
}

// sin(pi v), cos(pi v) for v in [0, 2): quadrant reduction to |x| <= 1/4, degree-15/14 Taylor in x (remainder 1e-15).
PCG_DEV void sincospi_02(double v, double& s, double& c) {
  const double q = rint(v * 2.0);                   // nearest half-integer count, 0..4
  const double x = fma(q, -0.5, v);                 // |x| <= 1/4 exactly representable reduction
  const int k = static_cast<int>(q);
  const double z = x * x;
  double ps = kMC.sn[7], pc = kMC.cs[7];
#pragma unroll
  for (int i = 6; i >= 0; --i) {
    ps = fma(ps, z, kMC.sn[i]);
    pc = fma(pc, z, kMC.cs[i]);
  }
  ps *= x;
  const bool swap = k & 1;
  const double ss = swap ? pc : ps, cc = swap ? ps : pc;
  s = (k & 2) ? -ss : ss;                           // k=0: (s,c)  1: (c,-s)  2: (-s,-c)  3: (-c,s)  4: (s,c)
  c = ((k + 1) & 2) ? -cc : cc;
}

}  // namespace pcg
