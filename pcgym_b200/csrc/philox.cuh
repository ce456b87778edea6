// Counter-based Gaussian draws for the measurement-noise block (reference pcgym.py:453-466
// uses the unseeded process-global np.random, so bitwise RNG parity is impossible by
// construction; the draw is a pure function of (seed, global env id, time index, pair index)
// instead, reproducible across launches, batch splits and GPUs).
//
// Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (env_lo, env_hi, t, pair).
#pragma once
#include <cstdint>

#include "fastmath.cuh"

namespace pcg {

// Round keys k_r = key + r * (W0, W1), computed once on the host (make_round_keys) and passed in the constant
// bank, so the ten rounds are two multiplies and two 3-input xors each.
struct PhiloxKeys { uint32_t k[10][2]; };

inline PhiloxKeys make_round_keys(uint64_t seed) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  PhiloxKeys rk;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    rk.k[r][0] = k0;
    rk.k[r][1] = k1;
    k0 += W0;
    k1 += W1;
  }
  return rk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& rk) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ rk.k[r][0], lo1, hi0 ^ c.w ^ rk.k[r][1], lo0);
  }
  return c;
}

// ---------------------------------------------------------------------------------------------------
// Draw layout (all kernels, all shardings): one Philox block serves the env PAIR (gid >> 1) at time t for the
// state pair j; env (gid & 1) takes words (x, y) or (z, w).  Two 32-bit words -> two N(0,1):
//     u0 = (w0 + 1/2) 2^-32,  u1 = (w1 + 1/2) 2^-32,   z0 = sqrt(-2 ln u0) cos(2 pi u1),  z1 = ... sin(2 pi u1)
// (32-bit uniforms: tails reach 6.7 sigma -- measurement noise; the arithmetic is fp64 to ~1e-15).

static __device__ const double2 kLogTab[128] = {
#include "logtab.inc"
};

struct BmConst {
  double m2ln2;       // -2 ln 2
  double lp[5];       // -2 ln(1+s) = -2 s + s^2 (lp0 + lp1 s + ... + lp4 s^4),  |s| <= 2^-8 (remainder 2/7 s^7 < 5e-18)
};
static __constant__ BmConst kBM = {-1.3862943611198906188, {1.0, -2.0 / 3, 0.5, -0.4, 1.0 / 3}};

// y = -2 ln((w + 1/2) 2^-32) > 0: exponent/7-bit mantissa bucket of d = 2w + 1, table of 1/c and -2 ln c, degree-6
// series in s = d_mant / c - 1.
PCG_DEV double neg2ln_u32(uint32_t w) {
  const double d = fma(static_cast<double>(w), 2.0, 1.0);          // odd 33-bit integer, exact; u = d 2^-33
  const int hi = __double2hiint(d);
  const int j = (hi >> 13) & 127;
  const double2 tab = __ldg(&kLogTab[j]);
  const int e = (hi >> 20) - (1023 + 33) + (j >= 53 ? 1 : 0);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(d));
  const double s = fma(m, tab.x, -1.0);
  double q = kBM.lp[4];
#pragma unroll
  for (int i = 3; i >= 0; --i) q = fma(q, s, kBM.lp[i]);
  const double base = fma(static_cast<double>(e), kBM.m2ln2, tab.y);
  return fma(s * s, q, fma(-2.0, s, base));
}

// sqrt(a), a > 0 normal: MUFU.RSQ64H seed (2^-23), two coupled Goldschmidt steps (g -> sqrt a, h -> 1/(2 sqrt a)).
PCG_DEV double sqrt_gs(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double g = a * r, h = 0.5 * r;
  double e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, h, 0.5);
  return fma(g, e, g);
}

// sin, cos of 2 pi (w + 1/2) 2^-32: quadrant q = nearest multiple of pi/2 and the signed remainder come straight
// from the integer, x = remainder / pi in [-1/4, 1/4]; same series as sincospi_02.
PCG_DEV void sincos_u32(uint32_t w, double& s, double& c) {
  const uint32_t q = ((w >> 29) + 1u) >> 1;                        // 0..4
  const int rem = static_cast<int>(w - (q << 30));                 // in [-2^29, 2^29), wraps correctly for q = 4
  const double x = fma(static_cast<double>(rem), 0x1p-31, 0x1p-32);
  const double z = x * x;
  double ps = kMC.sn[7], pc = kMC.cs[7];
#pragma unroll
  for (int i = 6; i >= 0; --i) {
    ps = fma(ps, z, kMC.sn[i]);
    pc = fma(pc, z, kMC.cs[i]);
  }
  ps *= x;
  const bool swap = q & 1u;
  const double ss = swap ? pc : ps, cc = swap ? ps : pc;
  // q=0: (s,c)  1: (c,-s)  2: (-s,-c)  3: (-c,s)  4: (s,c); signs applied in the integer pipe
  const int sneg = (q & 2u) ? 0x80000000 : 0, cneg = ((q + 1u) & 2u) ? 0x80000000 : 0;
  s = __hiloint2double(__double2hiint(ss) ^ sneg, __double2loint(ss));
  c = __hiloint2double(__double2hiint(cc) ^ cneg, __double2loint(cc));
}

template <bool F32>
PCG_DEV void normal_pair_u32(uint32_t w0, uint32_t w1, double& z0, double& z1) {
  if constexpr (F32) {                                             // Box-Muller evaluated in fp32 (24-bit uniforms)
    const float u0 = (static_cast<float>(w0 >> 8) + 0.5f) * 0x1p-24f;
    const float u1 = (static_cast<float>(w1 >> 8) + 0.5f) * 0x1p-24f;
    const float rad = sqrtf(-2.0f * logf(u0));
    float sn, cs;
    sincospif(2.0f * u1, &sn, &cs);
    z0 = static_cast<double>(rad * cs);
    z1 = static_cast<double>(rad * sn);
  } else {
    const double rad = sqrt_gs(neg2ln_u32(w0));
    double sn, cs;
    sincos_u32(w1, sn, cs);
    z0 = rad * cs;
    z1 = rad * sn;
  }
}

// N(0,1) draws for states (2j, 2j+1) of the EPT consecutive envs starting at global id gid0 at time index t.
// EPT == 2 requires an even gid0 (the launchers check env_id0): both envs then share one Philox block.
template <int EPT, bool F32>
PCG_DEV void draw_state_pair(uint64_t gid0, uint32_t t, uint32_t j, const PhiloxKeys& key, double (&z0)[EPT],
                             double (&z1)[EPT]) {
  const uint64_t pid = gid0 >> 1;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(pid), static_cast<uint32_t>(pid >> 32), t, j), key);
  if constexpr (EPT == 2) {
    normal_pair_u32<F32>(r.x, r.y, z0[0], z1[0]);
    normal_pair_u32<F32>(r.z, r.w, z0[1], z1[1]);
  } else {
    const bool odd = gid0 & 1u;
    normal_pair_u32<F32>(odd ? r.z : r.x, odd ? r.w : r.y, z0[0], z1[0]);
  }
}

}  // namespace pcg
