// Branch-free fp64 helpers for the RHS hot loops.
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
};

static __constant__ MathConst kMC = {
    1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 6755399441055744.0,
    {1.0, 1.0000000000000067, 0.5000000000000006, 0.16666666666554325, 0.04166666666657431, 0.008333333385696902,
     0.0013888888932304278, 0.0001984117023243583, 2.4801504458882045e-05, 2.764019693014132e-06,
     2.762633679135001e-07},
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

// n = rint(x log2 e) read back from the shifted double t = x log2 e + 1.5*2^52, clamped to +-1000.  The mantissa of t
// holds n as a 52-bit two's-complement integer; its low word alone wraps for |n| >= 2^31 (a hugely negative argument
// would come back with a positive n), so the high word decides: it equals the sign extension of the low word exactly
// when n fits 32 bits, and its own sign gives the saturation side otherwise.
PCG_DEV int exp_n_sat(double t) {
  const int lo = __double2loint(t), hd = __double2hiint(t) - 0x43380000;
  const int n = (hd == (lo >> 31)) ? lo : (hd < 0 ? -1000 : 1000);
  return max(-1000, min(1000, n));
}

// exp(x), saturating at 2^+-1000 instead of overflowing, NaN-propagating.  (Saturation holds while x log2 e fits the
// shifted double, |x| < ~1.5e9; further out the reduced argument is meaningless and the result is inf or NaN -- never a
// finite wrong value, so a state that far gone still reaches the non-finite status check.  test: pcg_math_probe.)
// n = rint(x log2 e) by the 1.5*2^52 shift, r = x - n ln2 (two-part ln2), degree-10 near-minimax polynomial
// in Estrin form (dependency depth 4 instead of Horner's 10), scaled by 2^n built in the exponent field.
PCG_DEV double exp_sat(double x) {
  const double t = fma(x, kMC.log2e, kMC.shift);
  const int n = exp_n_sat(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
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
  const int n = exp_n_sat(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
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
  const int n = exp_n_sat(t);
  const double nf = t - kMC.shift;
  double r = fma(nf, kMC.ln2_hi, x);
  r = fma(nf, kMC.ln2_lo, r);
  double p = kMC.ex[10];
#pragma unroll
  for (int i = 9; i >= 0; --i) p = fma(p, r, kMC.ex[i]);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// 1/a with ONE third-order correction of the MUFU.RCP64H seed (relative error e <= 2^-23):
// r (1 + e + e^2) = (1/a)(1 - e^3), i.e. 2^-69 before rounding -- three dependent DFMA instead of rcp_nr's four.
PCG_DEV double rcp_o3(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  return fma(fma(r, e, r), e, r);
}

// Table-driven exp for the integrator warps of the warp-specialised kernels (the serial part of the rollout, where
// both the FP64 instruction count and the LENGTH of the dependency chain set the pace):
//   n = rint(256 x log2 e) by the 1.5*2^52 shift, x = (n >> 8) ln2 + (n & 255) ln2/256 + r,  |r| <= ln2/512,
//   exp(x) = 2^(n >> 8) * 2^((n & 255)/256) * (1 + r + r^2/2 + r^3/6 + r^4/24)        (remainder r^5/120 < 4e-17)
// Eight FP64 instructions and a dependency depth of 7 against 14 / 13 for exp_sat_horner.  The reduction uses ln2/256
// as ONE constant: its rounding error times n is ~|x| 0.8e-16 relative in the result, the size of the rounding
// error x itself carries.  `tab` is a shared-memory copy of kExp2Tab.  Returned in two factors so the caller can
// fold its own multiplier into `scale` while the polynomial is still in flight: exp(x) = scale * poly.
// Saturates at 2^+-1000 for |256 x log2 e| < 2^39 (the shifted double holds n as a 52-bit two's-complement integer;
// bits 8..39 are used), NaN-propagating (poly is NaN).
struct ExpTabConst { double l256, nln2_256, shift, c2, c3, c4; };
static __constant__ ExpTabConst kET = {369.3299304675746, -0.0027076061740622863, 6755399441055744.0,
                                       0.5, 1.0 / 6, 1.0 / 24};
static __device__ const double kExp2Tab[256] = {
#include "exp2tab.inc"
};

PCG_DEV void exp_tab_split(double x, const double* tab, double& scale, double& poly) {
  const double t = fma(x, kET.l256, kET.shift);
  const int lo = __double2loint(t), hd = __double2hiint(t) - 0x43380000;
  const double nf = t - kET.shift;
  const double r = fma(nf, kET.nln2_256, x);
  int m = __funnelshift_r(lo, hd, 8);
  m = max(-1000, min(1000, m));
  const double tj = tab[lo & 255];
  scale = __hiloint2double(__double2hiint(tj) + (m << 20), __double2loint(tj));
  poly = fma(fma(fma(fma(kET.c4, r, kET.c3), r, kET.c2), r, 1.0), r, 1.0);
}

// The same exponential with the table indexed DIRECTLY by n = rint(256 x log2 e): tab[n - n_min] = 2^(n/256) for every
// n the caller's argument range can produce (the per-SM kernel builds it in shared memory from kExp2Tab at CTA start:
// exp_wide_fill).  `shift` = 1.5*2^52 - n_min puts n - n_min into the low word of the shifted double, so the whole
// integer side of the evaluation is ONE unsigned clamp and the address of a shared-memory load -- exp_tab_split's
// split / saturate / mask / exponent-add sequence (9 integer instructions per call, 54 per Tsit5 step, each an issue
// slot of the bound resource) is gone.  Same FP64 work, same rounding (2^(n/256) is the 256-entry
// table value with an exact power-of-two scaling).  An argument outside the table reads its first / last entry: the
// caller flags such lanes (PCG_ST_RANGE).  NaN-propagating through `poly`.
PCG_DEV void exp_wide_split(double x, const double* tab, double shift, unsigned last, double& scale, double& poly) {
  const double t = fma(x, kET.l256, shift);
  const unsigned idx = min(static_cast<unsigned>(__double2loint(t)), last);
  const double nf = t - shift;
  const double r = fma(nf, kET.nln2_256, x);
  scale = tab[idx];
  poly = fma(fma(fma(fma(kET.c4, r, kET.c3), r, kET.c2), r, 1.0), r, 1.0);
}

// whole CTA: tab[i] = 2^((i + n_min)/256), i = 0..last
PCG_DEV void exp_wide_fill(double* tab, int n_min, int last) {
  for (int i = threadIdx.x; i <= last; i += blockDim.x) {
    const int n = i + n_min;
    const double tj = kExp2Tab[n & 255];
    tab[i] = __hiloint2double(__double2hiint(tj) + ((n >> 8) << 20), __double2loint(tj));
  }
}

// sqrt(a) for the right-hand sides (four_tank: four per evaluation; the CUDA library's sqrt carries a slow-path branch
// that splits the basic block, 30 instructions each): MUFU.RSQ64H seed (2^-22), one coupled Goldschmidt step for
// (sqrt a, 1 / (2 sqrt a)) and a final residual correction -- 7 FP64 instructions, <= 1 ulp.  a < 0 and NaN give NaN like
// the reference's np.sqrt; a == 0 gives 0 (the seed is inf there: selected away); subnormal arguments are not supported.
PCG_DEV double sqrt_fast(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double g = a * r, h = 0.5 * r;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, a), h, g);
  return a == 0.0 ? a : g;
}

// log(x) for x > 0 (normal), branch-free: x = 2^e m, m in [1, 2); the top seven mantissa bits pick c_j with
// r = m / c_j - 1 formed EXACTLY by one fma against a 12-bit reciprocal (|r| < 2^-7.9), and
// log x = e ln2 + log c_j + (r - r^2/2 + r^3 (1/3 - r/4 + r^2/5 - r^3/6 + r^4/7))        (next term r^8/8 < 2e-20).
// ~1 ulp (tools/gen_logtab.py --check: 3.6e-15 absolute at |log| = 30, 2e-15 relative), 12 FP64 + 1 conversion + one
// 16-byte table load against ~45 instructions and three branches for the library's log.  x = 0 returns ~ -709 (its
// exponent field), x < 0 the log of |x| (crystallization wants exp(k log|x|), ~0 at x = 0).  NaN or inf in, NaN out.
struct LogTabConst { double ln2_hi, ln2_lo, c3, c4, c5, c6, c7; };
static __constant__ LogTabConst kLT = {0.6931471803691238, 1.9082149292705877e-10, 1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7};
static __device__ const double2 kLogTab[128] = {
#define PCG_LT(a, b) {a, b}
#include "logtab.inc"
#undef PCG_LT
};
PCG_DEV double log_tab(double x) {
  const int hi = __double2hiint(x);
  const int e = ((hi >> 20) & 0x7ff) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = __ldg(&kLogTab[(hi >> 13) & 127]);
  const double r = fma(m, t.x, -1.0);
  const double ed = static_cast<double>(e);
  double q = fma(kLT.c7, r, kLT.c6);
  q = fma(q, r, kLT.c5);
  q = fma(q, r, kLT.c4);
  q = fma(q, r, kLT.c3);
  const double p = fma(r * r, fma(r, q, -0.5), r);
  return fma(ed, kLT.ln2_hi, t.y) + fma(ed, kLT.ln2_lo, fma(x, 0.0, p));   // x * 0: NaN (and inf) in, NaN out
}

// high word of a double, host or device (positive normal doubles order like their high words)
PCG_HD int hi_word(double v) {
#ifdef __CUDA_ARCH__
  return __double2hiint(v);
#else
  long long b;
  static_assert(sizeof b == sizeof v, "");
  __builtin_memcpy(&b, &v, sizeof b);
  return static_cast<int>(b >> 32);
#endif
}

}  // namespace pcg
