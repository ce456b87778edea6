// Counter-based Gaussian draws for the measurement-noise block (reference pcgym.py:453-466
// uses the unseeded process-global np.random, so bitwise RNG parity is impossible by
// construction; the draw is a pure function of (seed, global env id, time index, pair index)
// instead, reproducible across launches, batch splits and GPUs).
//
// Philox4x32-10 (Salmon et al. 2011), key = seed (its high word also carries the episode counter, env.py),
// counter = (env_lo, env_hi, t, pair).
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
// state pair j; env (gid & 1) takes words (x, y) or (z, w); word k of the two is the draw of state 2j + k:
//     z = Phi^-1(u),  u = (w + 1/2) 2^-32          (inverse normal CDF of a 32-bit uniform: |z| <= 6.34)
// evaluated in fp64 by a table of piecewise degree-5 polynomials (csrc/icdftab.inc, tools/gen_icdf_tab.py):
//     fold     w >= 2^31: w' = ~w and the sign flips      p = (w' + 1/2) 2^-32 = m 2^-33,  m = 2 w' + 1
//     segment  b = leading-one index of m, sub = its next four bits           512 segments of 6 coefficients
//     local    v = the remaining bits of m, centred, as a signed 32-bit integer (coefficients are pre-scaled)
// 1 conversion + 5 DFMA + three 128-bit shared-memory loads per draw (the Box-Muller it replaced cost 36 FP64
// instructions per pair of draws); maximum absolute error 1.7e-12 against scipy.special.ndtri, where adjacent
// 32-bit uniforms are >= 5.8e-10 apart in z -- the transform is exact at the resolution of its input.
// `tab` is a shared-memory copy of kIcdfTab (kIcdfTab itself only in pcg_normal_pairs).

constexpr int kIcdfDoubles = 512 * 6;
static __device__ const double kIcdfTab[kIcdfDoubles] = {
#include "icdftab.inc"
};

// Every kernel family reads the table from SHARED memory: four 128-bit loads per draw with a different segment in
// every lane are ~32 L1 tag look-ups each from global memory (measured: the pair-per-CTA kernel went from 0.342 to
// 0.432 ms at 131 072 envs with the table in global memory), but only bank conflicts in shared memory.
PCG_DEV void fill_icdf_tab(double* tab) {          // whole CTA; the caller synchronises
  for (int i = threadIdx.x; i < kIcdfDoubles; i += blockDim.x) tab[i] = kIcdfTab[i];
}

PCG_DEV double normal_icdf_u32(uint32_t w, const double* tab) {
  const int s = static_cast<int>(w) >> 31;                             // all ones: upper half of the uniform
  const uint32_t m = ((w ^ static_cast<uint32_t>(s)) << 1) | 1u;
  const int lz = __clz(m);
  const uint32_t mn = m << lz;                                         // leading one at bit 31
  const uint32_t seg = static_cast<uint32_t>(31 - lz) * 16u + (mn >> 27);  // + 16: taken off the base below
  const int v = static_cast<int>((mn << 5) ^ 0x80000000u);
  const double x = static_cast<double>(v);
  const double2* c = reinterpret_cast<const double2*>(tab) + (static_cast<int>(seg) - 16) * 3;
  const double2 c01 = c[0], c23 = c[1], c45 = c[2];
  double z = fma(c45.y, x, c45.x);
  z = fma(z, x, c23.y);
  z = fma(z, x, c23.x);
  z = fma(z, x, c01.y);
  z = fma(z, x, c01.x);
  return __hiloint2double(__double2hiint(z) ^ (s & 0x80000000), __double2loint(z));
}

// fp32 evaluation (PCG_NOISE_PHILOX_F32): Box-Muller on 24-bit uniforms, two words -> two draws
PCG_DEV void normal_pair_f32(uint32_t w0, uint32_t w1, double& z0, double& z1) {
  const float u0 = (static_cast<float>(w0 >> 8) + 0.5f) * 0x1p-24f;
  const float u1 = (static_cast<float>(w1 >> 8) + 0.5f) * 0x1p-24f;
  const float rad = sqrtf(-2.0f * logf(u0));
  float sn, cs;
  sincospif(2.0f * u1, &sn, &cs);
  z0 = static_cast<double>(rad * cs);
  z1 = static_cast<double>(rad * sn);
}

template <bool F32>
PCG_DEV void normal_pair_u32(uint32_t w0, uint32_t w1, const double* tab, double& z0, double& z1) {
  if constexpr (F32) {
    normal_pair_f32(w0, w1, z0, z1);
  } else {
    z0 = normal_icdf_u32(w0, tab);
    z1 = normal_icdf_u32(w1, tab);
  }
}

// N(0,1) draws for states (2j, 2j+1) of the EPT consecutive envs starting at global id gid0 at time index t.
// EPT == 2 requires an even gid0 (the launchers check env_id0): both envs then share one Philox block.
template <int EPT, bool F32>
PCG_DEV void draw_state_pair(uint64_t gid0, uint32_t t, uint32_t j, const PhiloxKeys& key, const double* tab,
                             double (&z0)[EPT], double (&z1)[EPT]) {
  const uint64_t pid = gid0 >> 1;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(pid), static_cast<uint32_t>(pid >> 32), t, j), key);
  if constexpr (EPT == 2) {
    normal_pair_u32<F32>(r.x, r.y, tab, z0[0], z1[0]);
    normal_pair_u32<F32>(r.z, r.w, tab, z0[1], z1[1]);
  } else {
    const bool odd = gid0 & 1u;
    normal_pair_u32<F32>(odd ? r.z : r.x, odd ? r.w : r.y, tab, z0[0], z1[0]);
  }
}

}  // namespace pcg
