// Lane-split, lane-asynchronous form of the adaptive control-interval kernel for staged (banded) models.
//
// Why it exists (profiles/README.md, r2_extraction_adaptive): multistage_extraction under the adaptive pair is the one
// north-star workload that is FP64-bound by a wide margin (~490 right-hand sides per control interval), and the generic
// kernel runs it badly: one thread carries all ten states, six stage derivatives of each and the controller -- 255
// registers, a fully unrolled step of > 1 000 instructions that overflows the instruction cache (`no instruction` is its
// top stall), 6.9 warps per SM at 32 768 envs (10.7 % occupancy, issue slots 32 % busy) -- and every lane of a warp
// waits at the end of each control interval for the lane that needed the most steps.
//
// Here ONE ENV IS A GROUP OF LANES, one lane per stage of the column (M::LANES = 5 or 10: six or three envs per warp):
//   * a lane owns the two states of its stage (X_s, Y_s), their stage derivatives and its share of the error norm;
//     the right-hand side needs X_{s-1} from the lane on its left and Y_{s+1} from the lane on its right: two
//     64-bit shuffles per evaluation (M::rhs_lane); the error norm is a fixed-order tree sum over the group followed by a
//     broadcast, so every lane of a group takes bit-identical step decisions;
//   * 5.3x as many warps of ~100 registers and a step body of ~260 instructions: the schedulers have work and the
//     loop stays in the instruction cache;
//   * the control-interval boundary is NOT a synchronisation point: each group walks its own interval index k through a
//     flat loop "start interval | attempt one step | finish interval", so a warp executes the longest TOTAL number of
//     steps of its groups over the whole launch (which averages out) instead of the per-interval maximum every interval.
//     The step itself is executed by all 32 lanes together (full-mask shuffles: a partial, run-time mask makes the
//     compiler wrap every shuffle in a WARPSYNC / collective sequence, measured 5x the instructions); groups that have
//     finished, and the two spare lanes of a five-stage warp, run along on frozen data and commit nothing.  Only the
//     start / finish blocks are divergent, and they exchange data through shared memory, not shuffles.
// Semantics are those of rollout_kernel's adaptive path (same controller, same status bits, same outputs); only the order
// of the additions in the error norm differs, so step sequences agree with it to the controller's own tolerance.
// Eligibility is decided on the host (launch.cuh: lanes_ok); everything else takes rollout_kernel.
#pragma once
#include "rollout_fast.cuh"

namespace pcg {

constexpr int kLanesBlock = 128;   // 4 warps
#ifndef PCG_LANES_MINB
#define PCG_LANES_MINB 5            // resident CTAs per SM the register budget is cut for: 96 registers (4: 106 registers, +5 % time; 6: 80 registers with spills, +4 %)
#endif

template <class M, class = void>
struct has_lanes : std::false_type {};
template <class M>
struct has_lanes<M, std::enable_if_t<(M::LANES > 1)>> : std::true_type {};

template <class M, class TB, int NOISE>
__global__ void __launch_bounds__(kLanesBlock, PCG_LANES_MINB)
lanes_rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io,
                     const __grid_constant__ Derived<M, TB> dv, const int64_t n, const int t0, const int k_steps) {
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND, GS = M::LANES, EPW = 32 / GS, S = TB::S;
  static_assert(NX == 2 * GS, "one lane per stage, two states per lane");
  __shared__ __align__(16) double icdf_s[NOISE == PCG_NOISE_PHILOX ? kIcdfDoubles : 2];
  __shared__ double xs_all[kLanesBlock / 32][EPW][NX];          // end state of an interval, exchanged inside a group
  if constexpr (NOISE == PCG_NOISE_PHILOX) {
    fill_icdf_tab(icdf_s);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / GS, s = lane - grp * GS;
  const int64_t env_raw = (static_cast<int64_t>(blockIdx.x) * (kLanesBlock / 32) + warp) * EPW + grp;
  const bool valid = grp < EPW && env_raw < n;                   // spare lanes / envs past the end run along, store nothing
  const int64_t env = valid ? env_raw : 0;
  constexpr unsigned FULL = 0xffffffffu;
  const int gbase = grp * GS;
  const unsigned gmask = valid ? ((1u << GS) - 1u) << gbase : (1u << lane);
  // neighbours inside the group; the column's two ends read the feeds from the first spare lane of the warp, which
  // (like every lane without a live env) reads itself: its differences are exactly zero, never a stray NaN
  constexpr int kFeedLane = EPW * GS;
  static_assert(kFeedLane < 32, "the lane-split kernel needs a spare lane for the feeds");
  const int src_l = !valid ? lane : (s > 0 ? lane - 1 : kFeedLane);
  const int src_r = !valid ? lane : (s < GS - 1 ? lane + 1 : kFeedLane);
  double* xs = xs_all[warp][grp < EPW ? grp : 0];

  const int flags = d.flags;
  const bool norm_a = flags & PCG_F_NORMALISE_A, norm_o = flags & PCG_F_NORMALISE_O;
  const bool penal = flags & PCG_F_R_PENALTY, done_on_v = flags & PCG_F_DONE_ON_VIOLATION;
  const int nsp = d.nsp, n_con = d.n_con;
  const int nsp_obs = (flags & PCG_F_SP_HIDDEN) ? 0 : nsp, n_obs = NX + nsp_obs;
  const typename M::P& mp = dv.p;
  const PhiloxKeys& key = dv.rk;
  const double dt = d.dt, rtol = d.rtol, atol = d.atol;

  // this lane's two states; group-uniform bookkeeping is carried redundantly by every lane of the group
  double x[2] = {0.5, 0.5};
  if (valid) {
    x[0] = io.x[static_cast<int64_t>(2 * s) * n + env];
    x[1] = io.x[static_cast<int64_t>(2 * s + 1) * n + env];
  }
  bool done = (valid && io.done_state && !(flags & PCG_F_EPISODE_START)) ? (io.done_state[env] != 0) : false;
  unsigned status = 0;
  int n_rhs = 0;
  double h = (valid && io.h_carry) ? io.h_carry[env] : 0.0;      // step size carried from interval to interval

  // sum over the lanes of each group in a fixed order, broadcast from its first lane (identical in every lane of a
  // group); executed by the whole warp
  auto group_sum = [&](double v) {
#pragma unroll
    for (int dl = 1; dl < GS; dl <<= 1) {
      const double t = __shfl_down_sync(FULL, v, dl);
      if ((s & (2 * dl - 1)) == 0 && s + dl < GS) v += t;
    }
    return __shfl_sync(FULL, v, gbase);
  };
  // affine constraints over [x | uk] on the state in xs (every lane evaluates every row: ~14 FMA each, once per
  // interval, uniform constant-bank reads; lane c mod GS stores row c); returns "any g > 0"
  auto constraints = [&](const double (&uk)[NUK], double* gout) {
    bool viol = false;
#pragma unroll 1
    for (int c = 0; c < n_con; ++c) {
      double g = d.con_h[c];
#pragma unroll
      for (int i = 0; i < NX; ++i) g = fma(d.con_gx[c][i], xs[i], g);
#pragma unroll
      for (int i = 0; i < NUK; ++i) g = fma(d.con_gu[c][i], uk[i], g);
      viol = viol || (g > 0.0);                                  // NaN compares false, as np.any(g > 0)
      if (gout && (c % GS) == s) __stcs(gout + static_cast<int64_t>(c) * n, g);
    }
    return viol;
  };

  int k = 0, steps = 0;
  bool fin = !valid || k_steps <= 0;                             // all intervals of this group are done
  bool start = !fin, need_k0 = false, rejected = false;
  double t_in = 0.0, uk[NUK], k0[2];
#pragma unroll
  for (int i = 0; i < NUK; ++i) uk[i] = (i < NU) ? 1.0 : d.d_default[i - NU];
  k0[0] = k0[1] = 0.0;
  typename M::LaneC lc;                                            // coefficients of the current interval (zero: frozen lane)
  M::lane_prep(mp, uk, false, lc);
  if (lane == kFeedLane) M::lane_feed(uk, x);                      // the feeds, held as this lane's constant "state"

#pragma unroll 1
  while (__any_sync(FULL, !fin)) {
    if (start) {
      // ------------------------------------------------------------ start of control interval k (pcgym.py:372-420)
      const int t = t0 + k;
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const double a = __ldcs(io.a + (static_cast<int64_t>(k) * NU + j) * n + env);
        uk[j] = norm_a ? fma(a, d.a_scale[j], d.a_off[j]) : a;
      }
#pragma unroll
      for (int i = 0; i < ND; ++i) uk[NU + i] = d.d_default[i];
      M::lane_prep(mp, uk, true, lc);
      if (t == 0 && n_con > 0) {                                 // t == 0 pre-check records cons_info[:, 0, :]
        xs[2 * s] = x[0];
        xs[2 * s + 1] = x[1];
        __syncwarp(gmask);
        const bool v0 = constraints(uk, io.g0 ? io.g0 + env : nullptr);
        if (v0 && done_on_v) done = true;
        __syncwarp(gmask);
      }
      t_in = 0.0;
      h = (h > 0.0 && h < dt) ? h : dt;
      steps = 0;
      rejected = false;
      start = false;
      need_k0 = true;
    }
    // the derivative at the interval's start (the input changed, so the FSAL value is stale): evaluated by the whole warp
    // whenever some group needs it -- once per group and interval, i.e. in a few per cent of the iterations
    if (__any_sync(FULL, need_k0)) {
      double kk[2];
      M::rhs_lane(mp, lc, x, src_l, src_r, kk);
      if (need_k0) {
        k0[0] = kk[0];
        k0[1] = kk[1];
        ++n_rhs;
        need_k0 = false;
      }
    }

    // ---------------------------------------------------------------- one attempted step (rollout_kernel.cuh: advance_adaptive_aux)
    const bool capped = steps >= d.max_steps;
    const double left = dt - t_in;
    const bool last = (h >= left * (1.0 - 1e-12));
    const double hh = last ? left : h;
    const typename M::LaneC lh = {hh * lc.cl, hh * lc.cg, hh * lc.ql, hh * lc.qg};   // step folded in: the stages return h f
    double hk[S][2];
    hk[0][0] = hh * k0[0];
    hk[0][1] = hh * k0[1];
#pragma unroll
    for (int i = 1; i < S; ++i) {
      double xi[2];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        double acc = x[m];
#pragma unroll
        for (int j = 0; j < i; ++j)
          if (TB::a(i, j) != 0.0) acc = fma(dv.a[i][j], hk[j][m], acc);
        xi[m] = acc;
      }
      M::rhs_lane(mp, lh, xi, src_l, src_r, hk[i]);
    }
    double xn[2], kn[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      double acc = x[m];
#pragma unroll
      for (int i = 0; i < S; ++i)
        if (TB::b(i) != 0.0) acc = fma(dv.b[i], hk[i][m], acc);
      xn[m] = acc;
    }
    M::rhs_lane(mp, lc, xn, src_l, src_r, kn);                   // FSAL stage (unscaled: the next step has its own h)
    double sq = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      double err = dv.e[S] * (hh * kn[m]);
#pragma unroll
      for (int i = 0; i < S; ++i)
        if (TB::e(i) != 0.0) err = fma(dv.e[i], hk[i][m], err);
      const double sc = fma(rtol, fmax(fabs(x[m]), fabs(xn[m])), atol);
      const double r = err * rcp_nr(sc);
      sq = fma(r, r, sq);
    }
    const double msq = group_sum(sq) * (1.0 / NX);               // mean(err^2): accept test and step factor without a root
    const bool finite = isfinite(msq);
    const float grow = finite ? 0.9f * __powf(fmaxf(static_cast<float>(msq), 1e-30f), -0.1f) : 0.2f;

    bool finish = false;
    if (!fin) {
      if (capped) {
        status |= PCG_ST_STEPCAP;
        finish = true;
      } else {
        ++steps;
        n_rhs += S;
        if (finite && msq <= 1.0) {
          t_in = last ? dt : t_in + hh;
          x[0] = xn[0]; x[1] = xn[1];
          k0[0] = kn[0]; k0[1] = kn[1];
          const double hn = hh * static_cast<double>(fminf(rejected ? 1.0f : 10.0f, grow));
          h = last ? fmax(h, hn) : hn;                           // a truncated last step must not shrink the carried step
          rejected = false;
          finish = !(t_in < dt);
        } else {
          h = hh * static_cast<double>(fmaxf(0.2f, fminf(grow, 1.0f)));
          rejected = true;
          if (!(h > 1e-14 * dt)) {
            status |= PCG_ST_NONFINITE;
            finish = true;
          }
        }
      }
    }

    if (finish) {
      // -------------------------------------------------------------- end of control interval k (pcgym.py:432-498)
      const int t = t0 + k, t_new = t + 1;
      xs[2 * s] = x[0];
      xs[2 * s + 1] = x[1];
      __syncwarp(gmask);
      bool violated = false;
      if (n_con > 0) {
        violated = constraints(uk, io.g ? io.g + static_cast<int64_t>(k) * n_con * n + env : nullptr);
        if (violated && done_on_v) done = true;
      }
      if (t_new == d.N - 1) done = true;
      // setpoints: SP[t] is shown in the observation, SP[t+1] enters the reward (pcgym.py:432-441, 555)
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < kFastNsp; ++j)
        if (j < nsp) {
          const double* psp = io.sp + j * io.sp_sj + env * io.sp_se + static_cast<int64_t>(t) * io.sp_st;
          const double sp_cur = psp[0], sp_nxt = psp[io.sp_st];
          if (d.reward == PCG_REWARD_SP) {
            const double err = xs[d.sp_state[j]] - sp_nxt;
            r = fma(-(err * err), d.r_scale[j], r);
            if (penal && violated) r -= d.penalty;               // once per SP key
          }
          if (io.obs && j < nsp_obs && s == j) {
            const int i = NX + j;
            __stcs(io.obs + (static_cast<int64_t>(k) * n_obs + i) * n + env, norm_o ? fma(sp_cur, d.o_scale[i], d.o_off[i]) : sp_cur);
          }
        }
      // measurement noise on this lane's two states: the same Philox block / inverse-CDF draw as every other kernel
      double z[2] = {0.0, 0.0};
      if constexpr (NOISE == PCG_NOISE_PHILOX) {
        double za[1], zb[1];
        draw_state_pair<1, false>(static_cast<uint64_t>(d.env_id0 + env), static_cast<uint32_t>(t_new), s, key, icdf_s, za, zb);
        z[0] = za[0];
        z[1] = zb[0];
      } else if constexpr (NOISE == PCG_NOISE_EXTERNAL) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
          if (d.noise_pct[2 * s + m] != 0.0) z[m] = __ldcs(io.z + (static_cast<int64_t>(k) * NX + 2 * s + m) * n + env);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int i = 2 * s + m;
        if (io.x_traj) __stcs(io.x_traj + (static_cast<int64_t>(k) * NX + i) * n + env, x[m]);
        if (io.obs) {
          const double ox = fma(z[m] * x[m], d.noise_pct[i], x[m]);                      // pcgym.py:453-466
          double v = norm_o ? fma(ox, d.o_scale[i], d.o_off[i]) : ox;                    // pcgym.py:483-493
          if (d.obs_zero_mask >> i & 1u) v = 0.0;                                        // pcgym.py:495-498
          __stcs(io.obs + (static_cast<int64_t>(k) * n_obs + i) * n + env, v);
        }
      }
      if (s == 0) {
        const int64_t row = static_cast<int64_t>(k) * n + env;
        if (io.reward) __stcs(io.reward + row, r);
        if (io.done) io.done[row] = done ? 1 : 0;
        if (io.violated) io.violated[row] = violated ? 1 : 0;
      }
      __syncwarp(gmask);                                         // xs is rewritten by the next interval's end
      ++k;
      fin = k >= k_steps;
      start = !fin;
    }
  }
  if (!valid) return;

  // ---- carried state back to memory; non-finite states are flagged once per launch (NaN is sticky)
  io.x[static_cast<int64_t>(2 * s) * n + env] = x[0];
  io.x[static_cast<int64_t>(2 * s + 1) * n + env] = x[1];
  const bool bad = __ballot_sync(gmask, !isfinite(x[0]) || !isfinite(x[1])) != 0u;
  if (s == 0) {
    if (bad) status |= PCG_ST_NONFINITE;
    if (io.done_state) io.done_state[env] = done ? 1 : 0;
    if (io.status && status) io.status[env] |= static_cast<uint8_t>(status);
    if (io.h_carry) io.h_carry[env] = h;
    if (io.n_rhs) io.n_rhs[env] += n_rhs;
  }
}

}  // namespace pcg
