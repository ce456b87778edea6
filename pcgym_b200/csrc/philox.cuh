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

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// two independent N(0,1) from one Philox block, Box-Muller in fp64.
// u in (0,1): 53 random bits + half an ulp, so log never sees 0.
__device__ __forceinline__ void normal_pair_f64(uint4 r, double& z0, double& z1) {
  const uint64_t b0 = (static_cast<uint64_t>(r.y) << 32) | r.x;
  const uint64_t b1 = (static_cast<uint64_t>(r.w) << 32) | r.z;
  const double u0 = (static_cast<double>(b0 >> 11) + 0.5) * 0x1p-53;
  const double u1 = (static_cast<double>(b1 >> 11) + 0.5) * 0x1p-53;
  const double rad = sqrt_nr(-2.0 * log_pos(u0));      // branch-free: the two envs of a thread interleave
  double s, c;
  sincospi_02(2.0 * u1, s, c);
  z0 = rad * c;
  z1 = rad * s;
}

// four N(0,1) from one Philox block, Box-Muller evaluated in fp32 (24-bit uniforms).
__device__ __forceinline__ void normal_quad_f32(uint4 r, double (&z)[4]) {
  const float u0 = (static_cast<float>(r.x >> 8) + 0.5f) * 0x1p-24f;
  const float u1 = (static_cast<float>(r.y >> 8) + 0.5f) * 0x1p-24f;
  const float u2 = (static_cast<float>(r.z >> 8) + 0.5f) * 0x1p-24f;
  const float u3 = (static_cast<float>(r.w >> 8) + 0.5f) * 0x1p-24f;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float sa, ca, sb, cb;
  sincospif(2.0f * u1, &sa, &ca);
  sincospif(2.0f * u3, &sb, &cb);
  z[0] = static_cast<double>(ra * ca);
  z[1] = static_cast<double>(ra * sa);
  z[2] = static_cast<double>(rb * cb);
  z[3] = static_cast<double>(rb * sb);
}

}  // namespace pcg
