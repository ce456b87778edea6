// The fused control-interval kernel: one thread (warp lane) per environment.
//
// For each of k_steps control intervals it restates make_env.step (reference
// src/pcgym/pcgym.py:350-500) for its environment:
//   un-normalise action (+ a_delta)  ->  assemble uk with disturbances  ->  [t==0 constraint pre-check]
//   ->  integrate the model over dt (zero-order hold)  ->  setpoint slots  ->  t += 1
//   ->  constraint values / violation flag / done  ->  Gaussian measurement noise
//   ->  setpoint-tracking (or terminal batch) reward  ->  observation normalisation + mask
// with the model state held in registers across intervals.  All global traffic is SoA
// [field][env]: a warp touches 32 consecutive doubles (two full 128-byte lines) per field.
//
// Loads of step k+1's inputs are issued before step k's integration (register double buffer),
// outputs are written with streaming stores (they are never re-read by this kernel).
#pragma once
#include <cstdint>

#include "../../include/pcgym_b200.h"
#include "models.cuh"
#include "philox.cuh"
#include "tableaus.cuh"

namespace pcg {

constexpr int kBlock = 64;   // 65 536 envs -> 1024 CTAs, 6.9 per SM: all resident, <2 % imbalance

template <class M>
struct Dims {
  static constexpr int NX = M::NX;
  static constexpr int NUK = M::NU + M::ND;     // uk length with disturbance channels
};

PCG_DEV double ld_stream(const double* p) { return __ldcs(p); }
PCG_DEV void st_stream(double* p, double v) { __stcs(p, v); }

template <int N>
PCG_DEV double pick(const double (&v)[N], int idx) {
  double r = v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
  return r;
}

// ------------------------------------------------------------------ explicit RK machinery
// stage values x_i = x + sum_j a_ij (h k_j); hk[j] holds h*k_j so the sums are pure FMAs
template <class M, class TB>
PCG_DEV void rk_stages(const typename M::P& p, const double (&x)[M::NX], const double (&u)[Dims<M>::NUK],
                       double h, double (&hk)[TB::S + 1][M::NX], bool have_k0) {
  constexpr int NX = M::NX;
  if (!have_k0) {
    double k[NX];
    M::rhs(p, x, u, k);
#pragma unroll
    for (int n = 0; n < NX; ++n) hk[0][n] = h * k[n];
  }
#pragma unroll
  for (int i = 1; i < TB::S; ++i) {
    double xi[NX], k[NX];
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      double acc = x[n];
#pragma unroll
      for (int j = 0; j < i; ++j)
        if (TB::a(i, j) != 0.0) acc = fma(TB::a(i, j), hk[j][n], acc);
      xi[n] = acc;
    }
    M::rhs(p, xi, u, k);
#pragma unroll
    for (int n = 0; n < NX; ++n) hk[i][n] = h * k[n];
  }
}

template <class M, class TB>
PCG_DEV void rk_solution(const double (&x)[M::NX], const double (&hk)[TB::S + 1][M::NX], double (&xn)[M::NX]) {
#pragma unroll
  for (int n = 0; n < M::NX; ++n) {
    double acc = x[n];
#pragma unroll
    for (int i = 0; i < TB::S; ++i)
      if (TB::b(i) != 0.0) acc = fma(TB::b(i), hk[i][n], acc);
    xn[n] = acc;
  }
}

// fixed step: n_sub steps of h = dt / n_sub
template <class M, class TB>
PCG_DEV void advance_fixed(const typename M::P& p, double (&x)[M::NX], const double (&u)[Dims<M>::NUK], double dt,
                           int n_sub, int& n_rhs) {
  const double h = dt / static_cast<double>(n_sub);
#pragma unroll 1
  for (int s = 0; s < n_sub; ++s) {
    double hk[TB::S + 1][M::NX];
    rk_stages<M, TB>(p, x, u, h, hk, false);
    double xn[M::NX];
    rk_solution<M, TB>(x, hk, xn);
#pragma unroll
    for (int n = 0; n < M::NX; ++n) x[n] = xn[n];
  }
  n_rhs += n_sub * TB::S;
}

// adaptive step: embedded 5(4) pair, per-lane error control.  Controller as in
// Hairer/Norsett/Wanner II.4 (and scipy RK45): err = RMS(e_i / (atol + rtol max(|x_i|,|x_i'|))),
// accept if err <= 1, h *= clamp(0.9 err^-1/5, 0.2, 10) (no growth right after a rejection).
// Lanes of a warp take different step counts; `max_steps` bounds the loop and a non-finite
// error estimate is treated as a rejection so one bad lane cannot stall its warp.
template <class M, class TB>
PCG_DEV void advance_adaptive(const typename M::P& p, double (&x)[M::NX], const double (&u)[Dims<M>::NUK], double dt,
                              double rtol, double atol, int max_steps, double& h_carry, int& n_rhs,
                              unsigned& status) {
  constexpr int NX = M::NX;
  double t = 0.0;
  double h = (h_carry > 0.0 && h_carry < dt) ? h_carry : dt;
  double k0[NX];
  M::rhs(p, x, u, k0);
  ++n_rhs;
  int steps = 0;
  bool rejected = false;
#pragma unroll 1
  while (t < dt) {
    if (steps >= max_steps) { status |= PCG_ST_STEPCAP; break; }
    ++steps;
    const double left = dt - t;
    const bool last = (h >= left * (1.0 - 1e-12));
    const double hh = last ? left : h;
    double hk[TB::S + 1][NX];
#pragma unroll
    for (int n = 0; n < NX; ++n) hk[0][n] = hh * k0[n];
    rk_stages<M, TB>(p, x, u, hh, hk, true);
    double xn[NX], kn[NX];
    rk_solution<M, TB>(x, hk, xn);
    M::rhs(p, xn, u, kn);                       // FSAL stage: k_{S+1} = f(x_new)
    n_rhs += TB::S;
    double sq = 0.0;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      double e = TB::e(TB::S) * (hh * kn[n]);
#pragma unroll
      for (int i = 0; i < TB::S; ++i)
        if (TB::e(i) != 0.0) e = fma(TB::e(i), hk[i][n], e);
      const double sc = fma(rtol, fmax(fabs(x[n]), fabs(xn[n])), atol);
      const double r = e / sc;
      sq = fma(r, r, sq);
    }
    const double err = sqrt(sq * (1.0 / NX));
    const bool finite = isfinite(err);
    const float grow = finite ? 0.9f * __powf(fmaxf(static_cast<float>(err), 1e-30f), -0.2f) : 0.2f;
    if (finite && err <= 1.0) {
      t = last ? dt : t + hh;
#pragma unroll
      for (int n = 0; n < NX; ++n) { x[n] = xn[n]; k0[n] = kn[n]; }
      const float f = fminf(rejected ? 1.0f : 10.0f, grow);
      const double hn = hh * static_cast<double>(f);
      h = last ? fmax(h, hn) : hn;             // a truncated last step must not shrink the carried step
      rejected = false;
    } else {
      h = hh * static_cast<double>(fmaxf(0.2f, fminf(grow, 1.0f)));
      rejected = true;
      if (!(h > 1e-14 * dt)) { status |= PCG_ST_NONFINITE; break; }
    }
  }
  h_carry = h;
}

// ------------------------------------------------------------------ the kernel
template <class M, class TB, bool ADAPTIVE>
__global__ void __launch_bounds__(kBlock)
rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io, const int64_t n_env,
               const int t0, const int k_steps) {
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND;
  const int64_t env = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
  if (env >= n_env) return;
  const int64_t n = n_env;

  const typename M::P p = M::make(d.theta);
  const int flags = d.flags;
  const bool norm_a = flags & PCG_F_NORMALISE_A;
  const bool norm_o = flags & PCG_F_NORMALISE_O;
  const bool a_delta = norm_a && (flags & PCG_F_A_DELTA) && io.a_save != nullptr;   // pcgym.py:376
  const bool dist_on = d.nd > 0;
  const int nsp = d.nsp, n_con = d.n_con, nd_obs = d.nd_obs, n_tail = d.n_tail;
  const int n_obs = NX + nsp + nd_obs + n_tail;

  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = io.x[i * n + env];
  double a_save[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) a_save[j] = a_delta ? io.a_save[j * n + env] : 0.0;
  double tail[PCG_MAX_NTAIL];
#pragma unroll
  for (int j = 0; j < PCG_MAX_NTAIL; ++j) tail[j] = (j < n_tail && io.tail) ? io.tail[j * n + env] : 0.0;

  bool done = io.done_state ? (io.done_state[env] != 0) : false;
  unsigned status = 0;
  int n_rhs = 0;
  double h_carry = (ADAPTIVE && io.h_carry) ? io.h_carry[env] : 0.0;

  // register double buffer for the streamed inputs
  double a_nxt[NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) a_nxt[j] = ld_stream(io.a + (static_cast<int64_t>(0) * NU + j) * n + env);

#pragma unroll 1
  for (int k = 0; k < k_steps; ++k) {
    const int t = t0 + k;
    double a_raw[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) a_raw[j] = a_nxt[j];
    if (k + 1 < k_steps) {
#pragma unroll
      for (int j = 0; j < NU; ++j) a_nxt[j] = ld_stream(io.a + (static_cast<int64_t>(k + 1) * NU + j) * n + env);
    }

    // ---- action -> physical input (pcgym.py:372-383)
    double uk[NUK];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double a = a_raw[j];
      if (norm_a) a = fma(a, d.a_scale[j], d.a_off[j]);
      if (a_delta) {
        a = fma(a, d.a_scale[j], d.a_off[j]);       // normalised a second time, as the reference does
        a = a_save[j] + a;                          // the UNclipped sum is integrated ...
        a_save[j] = fmin(fmax(a, d.a_act_lo[j]), d.a_act_hi[j]);   // ... only a_save is clipped
      }
      uk[j] = a;
    }
    // ---- disturbances at t+1 (pcgym.py:386-412); defaults when inactive so the RHS needs no branch
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      double v = d.d_default[i];
      if (dist_on && d.d_src[i] >= 0)
        v = io.dist[static_cast<int64_t>(t + 1) * io.d_st + d.d_src[i] * io.d_sj + env * io.d_se];
      uk[NU + i] = v;
    }
    double dslot[PCG_MAX_ND];                        // observed disturbance slots of the state vector
#pragma unroll
    for (int j = 0; j < PCG_MAX_ND; ++j) {
      double v = 0.0;
      if (ND > 0 && j < nd_obs) {
#pragma unroll
        for (int i = 0; i < ND; ++i) v = (d.d_obs[j] == i) ? uk[NU + i] : v;
      }
      dslot[j] = v;
    }

    // ---- constraint evaluator over state' = [x | sp | dslot | tail] and uk
    auto constraints = [&](const double (&xs)[NX], const double (&sps)[PCG_MAX_NSP], double* gout, int64_t gstride) {
      bool viol = false;
#pragma unroll 1
      for (int c = 0; c < n_con; ++c) {
        double g = d.con_h[c];
        const uint32_t mx = d.con_mask_x[c], mu = d.con_mask_u[c];
#pragma unroll
        for (int i = 0; i < NX; ++i)
          if (mx >> i & 1u) g = fma(d.con_gx[c][i], xs[i], g);
        if (mx >> NX) {                               // rare: constraint reads SP / disturbance / tail slots
#pragma unroll
          for (int j = 0; j < PCG_MAX_NSP; ++j)
            if (j < nsp && (mx >> (NX + j) & 1u)) g = fma(d.con_gx[c][NX + j], sps[j], g);
#pragma unroll
          for (int j = 0; j < PCG_MAX_ND; ++j)
            if (j < nd_obs && (mx >> (NX + nsp + j) & 1u)) g = fma(d.con_gx[c][NX + nsp + j], dslot[j], g);
#pragma unroll
          for (int j = 0; j < PCG_MAX_NTAIL; ++j)
            if (j < n_tail && (mx >> (NX + nsp + nd_obs + j) & 1u))
              g = fma(d.con_gx[c][NX + nsp + nd_obs + j], tail[j], g);
        }
#pragma unroll
        for (int i = 0; i < NUK; ++i)
          if (mu >> i & 1u) g = fma(d.con_gu[c][i], uk[i], g);
        if (gout) st_stream(gout + c * gstride, g);
        viol = viol || (g > 0.0);                     // NaN compares false, as np.any(g > 0)
      }
      return viol;
    };

    // ---- t == 0 pre-check (pcgym.py:416-420): records cons_info[:,0,:], may set done
    if (t == 0 && n_con > 0) {
      double sp0[PCG_MAX_NSP];
#pragma unroll
      for (int j = 0; j < PCG_MAX_NSP; ++j) sp0[j] = d.sp_init[j];
      const bool v0 = constraints(x, sp0, io.g0 ? io.g0 + env : nullptr, n);
      if (v0 && (flags & PCG_F_DONE_ON_VIOLATION)) done = true;
    }

    // ---- integrate over the control interval, input held (integrator.py:90-107)
    if constexpr (ADAPTIVE && TB::EMBEDDED) {
      advance_adaptive<M, TB>(p, x, uk, d.dt, d.rtol, d.atol, d.max_steps, h_carry, n_rhs, status);
    } else {
      advance_fixed<M, TB>(p, x, uk, d.dt, d.n_sub, n_rhs);
    }
    bool finite = true;
#pragma unroll
    for (int i = 0; i < NX; ++i) finite = finite && isfinite(x[i]);
    if (!finite) status |= PCG_ST_NONFINITE;

    // ---- setpoint slots carry SP[t] (pcgym.py:432-438); the reward below uses SP[t+1]
    double sp_obs[PCG_MAX_NSP], sp_rew[PCG_MAX_NSP];
#pragma unroll
    for (int j = 0; j < PCG_MAX_NSP; ++j) {
      sp_obs[j] = sp_rew[j] = 0.0;
      if (j < nsp) {
        const double* base = io.sp + j * io.sp_sj + env * io.sp_se;
        sp_obs[j] = base[static_cast<int64_t>(t) * io.sp_st];
        sp_rew[j] = base[static_cast<int64_t>(t + 1) * io.sp_st];
      }
    }
    const int t_new = t + 1;                                       // pcgym.py:441

    // ---- constraints on the new state (pcgym.py:444-446)
    bool violated = false;
    if (n_con > 0) {
      violated = constraints(x, sp_obs, io.g ? io.g + static_cast<int64_t>(k) * n_con * n + env : nullptr, n);
      if (violated && (flags & PCG_F_DONE_ON_VIOLATION)) done = true;
    }
    if (t_new == d.N - 1) done = true;                             // pcgym.py:448-449

    // ---- reward on the noise-free state (pcgym.py:470-482)
    double r = 0.0;
    if (d.reward == PCG_REWARD_SP) {                               // pcgym.py:535-558
#pragma unroll
      for (int j = 0; j < PCG_MAX_NSP; ++j) {
        if (j < nsp) {
          const double e = pick<NX>(x, d.sp_state[j]) - sp_rew[j];
          r = fma(-(e * e), d.r_scale[j], r);
          if ((flags & PCG_F_R_PENALTY) && violated) r -= d.penalty;   // once per SP key
        }
      }
    } else if (d.reward == PCG_REWARD_BATCH) {                     // pcgym.py:502-532
      if (t_new == d.N - 1) {
#pragma unroll
        for (int i = 0; i < NX; ++i) r = fma(d.batch_w[i], x[i], r);
        if ((flags & PCG_F_R_PENALTY) && violated) r -= d.penalty;
      }
    }

    // ---- measurement noise on the model states only (pcgym.py:453-466)
    double ox[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) ox[i] = x[i];
    if (d.noise == PCG_NOISE_EXTERNAL) {
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if (d.noise_pct[i] != 0.0) {
          const double z = ld_stream(io.z + (static_cast<int64_t>(k) * NX + i) * n + env);
          ox[i] = fma(z * x[i], d.noise_pct[i], x[i]);
        }
    } else if (d.noise == PCG_NOISE_PHILOX) {
      const uint64_t gid = static_cast<uint64_t>(d.env_id0 + env);
      const uint2 key = make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
#pragma unroll
      for (int i = 0; i < NX; i += 2) {
        double z0, z1;
        normal_pair_f64(philox4x32_10(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                                 static_cast<uint32_t>(t_new), i / 2), key), z0, z1);
        ox[i] = fma(z0 * x[i], d.noise_pct[i], x[i]);
        if (i + 1 < NX) ox[i + 1] = fma(z1 * x[i + 1], d.noise_pct[i + 1], x[i + 1]);
      }
    } else if (d.noise == PCG_NOISE_PHILOX_F32) {
      const uint64_t gid = static_cast<uint64_t>(d.env_id0 + env);
      const uint2 key = make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
#pragma unroll
      for (int i = 0; i < NX; i += 4) {
        double z[4];
        normal_quad_f32(philox4x32_10(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32),
                                                 static_cast<uint32_t>(t_new), i / 4), key), z);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i + q < NX) ox[i + q] = fma(z[q] * x[i + q], d.noise_pct[i + q], x[i + q]);
      }
    }

    // ---- outputs
    const int64_t row = static_cast<int64_t>(k);
    if (io.x_traj) {
#pragma unroll
      for (int i = 0; i < NX; ++i) st_stream(io.x_traj + (row * NX + i) * n + env, x[i]);
    }
    if (io.obs) {
      double* o = io.obs + row * n_obs * n + env;
      const uint32_t zmask = d.obs_zero_mask;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = norm_o ? fma(ox[i], d.o_scale[i], d.o_off[i]) : ox[i];   // pcgym.py:483-493
        if (zmask >> i & 1u) v = 0.0;                                       // pcgym.py:495-498
        st_stream(o + i * n, v);
      }
#pragma unroll
      for (int j = 0; j < PCG_MAX_NSP; ++j)
        if (j < nsp) {
          const int i = NX + j;
          st_stream(o + i * n, norm_o ? fma(sp_obs[j], d.o_scale[i], d.o_off[i]) : sp_obs[j]);
        }
#pragma unroll
      for (int j = 0; j < PCG_MAX_ND; ++j)
        if (j < nd_obs) {
          const int i = NX + nsp + j;
          st_stream(o + i * n, norm_o ? fma(dslot[j], d.o_scale[i], d.o_off[i]) : dslot[j]);
        }
#pragma unroll
      for (int j = 0; j < PCG_MAX_NTAIL; ++j)
        if (j < n_tail) {
          const int i = NX + nsp + nd_obs + j;
          st_stream(o + i * n, norm_o ? fma(tail[j], d.o_scale[i], d.o_off[i]) : tail[j]);
        }
    }
    if (io.reward) st_stream(io.reward + row * n + env, r);
    if (io.done) io.done[row * n + env] = done ? 1 : 0;
    if (io.violated) io.violated[row * n + env] = violated ? 1 : 0;
  }

  // ---- carried state
#pragma unroll
  for (int i = 0; i < NX; ++i) io.x[i * n + env] = x[i];
  if (a_delta) {
#pragma unroll
    for (int j = 0; j < NU; ++j) io.a_save[j * n + env] = a_save[j];
  }
  if (io.done_state) io.done_state[env] = done ? 1 : 0;
  if (io.status && status) io.status[env] |= static_cast<uint8_t>(status);
  if (ADAPTIVE && io.h_carry) io.h_carry[env] = h_carry;
  if (io.n_rhs) io.n_rhs[env] += n_rhs;
}

}  // namespace pcg
