// The fused control-interval kernel.
//
// For each of k_steps control intervals a thread restates make_env.step (reference
// src/pcgym/pcgym.py:350-500) for the EPT consecutive environments it owns:
//   un-normalise action (+ a_delta)  ->  assemble uk with disturbances  ->  [t==0 constraint pre-check]
//   ->  integrate the model over dt (zero-order hold)  ->  setpoint slots  ->  t += 1
//   ->  constraint values / violation flag / done  ->  Gaussian measurement noise
//   ->  setpoint-tracking (or terminal batch) reward  ->  observation normalisation + mask
// with the model state held in registers across intervals.
//
// Performance notes (profiles/):
//   * The kernel is bound by the FP64 pipe, not HBM: ~300 DFMA-class ops against 74 B per CSTR env-step.
//     EPT = 2 environments per thread give the scheduler two independent dependency chains (measured
//     +25 % DFMA throughput at this occupancy, tools/fp64_microbench.cu) and turn every global access
//     into a 128-bit one (lane L owns envs 2L, 2L+1: a warp touches 512 contiguous bytes per row).
//   * Everything uniform lives in the constant bank: the descriptor, the model's derived constants and
//     the h-scaled Butcher coefficients are kernel parameters (one LDCU / LDC per use into a uniform or vector
//     register; sm_100 has no constant-bank operands on FP64 instructions), never recomputed per thread.
//   * GENERAL = false drops the rarely used observation slots (observed disturbances, uncertain-parameter
//     tail, > 2 setpoints, constraints that read them) from the instruction stream.
//   * Streamed inputs of step k+1 are loaded before step k is integrated; outputs use evict-first stores.
#pragma once
#include <cstdint>

#include "../../include/pcgym_b200.h"
#include "models_ext.cuh"
#include "philox.cuh"
#include "tableaus.cuh"

namespace pcg {

constexpr int kBlock = 64;   // 65 536 envs, EPT 2 -> 512 CTAs of 2 warps: 3.46 CTAs per SM, all resident

// Host-computed, passed by value (constant bank): model constants + Butcher coefficients.
template <class M, class TB>
struct Derived {
  typename M::P p;
  double ha[TB::S][TB::S];   // h * a_ij  (fixed mode, h = dt / n_sub)
  double hb[TB::S];          // h * b_i
  double a[TB::S][TB::S];    // raw coefficients (adaptive mode, h varies per lane)
  double b[TB::S];
  double e[TB::S + 1];
  PhiloxKeys rk;             // Philox round keys of desc.seed
};

template <class M, class TB>
inline Derived<M, TB> make_derived(const pcg_desc& d) {
  Derived<M, TB> v{};
  v.p = M::make(d.theta);
  const double h = d.dt / static_cast<double>(d.n_sub > 0 ? d.n_sub : 1);
  for (int i = 0; i < TB::S; ++i) {
    for (int j = 0; j < TB::S; ++j) {
      v.a[i][j] = (j < i) ? TB::a(i, j) : 0.0;
      v.ha[i][j] = h * v.a[i][j];
    }
    v.b[i] = TB::b(i);
    v.hb[i] = h * v.b[i];
  }
  for (int i = 0; i <= TB::S; ++i) v.e[i] = TB::e(i);
  v.rk = make_round_keys(d.seed);
  return v;
}

template <int N>
PCG_DEV double pick(const double (&v)[N], int idx) {
  double r = v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
  return r;
}

// ------------------------------------------------------------------ row access: EPT consecutive envs
// EPT == 2 is only launched when every row is 16-byte aligned (even n_env, aligned bases: launch.cuh),
// so the pair is one 128-bit access; EPT == 1 is the general 64-bit path.
template <int EPT>
PCG_DEV void ld_row(const double* p, double (&v)[EPT]) {
  if constexpr (EPT == 2) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[EPT - 1] = t.y;
  } else {
    v[0] = __ldcs(p);
  }
}
template <int EPT>
PCG_DEV void st_row(double* p, const double (&v)[EPT]) {
  if constexpr (EPT == 2) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[EPT - 1]));
  } else {
    __stcs(p, v[0]);
  }
}
template <int EPT>
PCG_DEV void st_flags(uint8_t* p, const bool (&f)[EPT]) {
  if constexpr (EPT == 2) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2(f[0] ? 1 : 0, f[EPT - 1] ? 1 : 0);
  } else {
    p[0] = f[0] ? 1 : 0;
  }
}

// Model constants as seen by EPT environments: the same constant-bank block for all of them (UniformP), or a
// per-environment copy built in registers from io.theta (an array of M::P).  Both index as pa[e].
template <class P>
struct UniformP {
  const P& p;
  PCG_DEV const P& operator[](int) const { return p; }
};

// ------------------------------------------------------------------ fixed step: n_sub steps of h = dt / n_sub
// The EPT environments advance in lock step: their RHS evaluations sit in one basic block so the
// scheduler interleaves the two dependency chains.
template <class M, class TB, int EPT, class PA>
PCG_DEV void rk_step_fixed(const Derived<M, TB>& dv, const PA& pa, const typename M::Aux (&aux)[EPT],
                           double (&x)[EPT][M::NX], const double (&u)[EPT][M::NU + M::ND]) {
  constexpr int NX = M::NX;
  {
    double k[TB::S][EPT][NX];
#pragma unroll
    for (int e = 0; e < EPT; ++e) M::rhs(pa[e], aux[e], x[e], u[e], k[0][e]);
#pragma unroll
    for (int i = 1; i < TB::S; ++i) {
      double xi[EPT][NX];
#pragma unroll
      for (int e = 0; e < EPT; ++e)
#pragma unroll
        for (int n = 0; n < NX; ++n) {
          double acc = x[e][n];
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (TB::a(i, j) != 0.0) acc = fma(dv.ha[i][j], k[j][e][n], acc);
          xi[e][n] = acc;
        }
#pragma unroll
      for (int e = 0; e < EPT; ++e) M::rhs(pa[e], aux[e], xi[e], u[e], k[i][e]);
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int n = 0; n < NX; ++n) {
        double acc = x[e][n];
#pragma unroll
        for (int i = 0; i < TB::S; ++i)
          if (TB::b(i) != 0.0) acc = fma(dv.hb[i], k[i][e][n], acc);
        x[e][n] = acc;
      }
  }
}
template <class M, class TB, int EPT>
PCG_DEV void rk_step_fixed(const Derived<M, TB>& dv, const typename M::Aux (&aux)[EPT], double (&x)[EPT][M::NX],
                           const double (&u)[EPT][M::NU + M::ND]) {
  rk_step_fixed<M, TB, EPT>(dv, UniformP<typename M::P>{dv.p}, aux, x, u);
}

// ------------------------------------------------------------------ adaptive step, one environment
// Embedded 5(4) pair, per-lane error control.  Controller as in Hairer/Norsett/Wanner II.4 (and scipy
// RK45): err = RMS(e_i / (atol + rtol max(|x_i|,|x_i'|))), accept if err <= 1,
// h *= clamp(0.9 err^-1/5, 0.2, 10) (no growth right after a rejection).  Lanes of a warp take
// different step counts; `max_steps` bounds the loop and a non-finite error estimate is treated as a
// rejection so one bad lane cannot stall its warp.
template <class M, class TB>
PCG_DEV void advance_adaptive_aux(const Derived<M, TB>& dv, const typename M::P& mp, const typename M::Aux& aux,
                                  double (&x)[M::NX], const double (&u)[M::NU + M::ND], double dt,
                                  double rtol, double atol, int max_steps, double& h_carry, int& n_rhs,
                                  unsigned& status);

template <class M, class TB>
PCG_DEV void advance_adaptive(const Derived<M, TB>& dv, const typename M::P& mp, double (&x)[M::NX],
                              const double (&u)[M::NU + M::ND], double dt,
                              double rtol, double atol, int max_steps, double& h_carry, int& n_rhs,
                              unsigned& status) {
  typename M::Aux aux;
  M::prep(mp, u, aux);
  advance_adaptive_aux<M, TB>(dv, mp, aux, x, u, dt, rtol, atol, max_steps, h_carry, n_rhs, status);
}

// the same with the per-interval constants already prepared (the warp-specialised kernels bind their shared-memory
// tables into `aux` first)
template <class M, class TB>
PCG_DEV void advance_adaptive_aux(const Derived<M, TB>& dv, const typename M::P& mp, const typename M::Aux& aux,
                                  double (&x)[M::NX], const double (&u)[M::NU + M::ND], double dt,
                                  double rtol, double atol, int max_steps, double& h_carry, int& n_rhs,
                                  unsigned& status) {
  constexpr int NX = M::NX, S = TB::S;
  double t = 0.0;
  double h = (h_carry > 0.0 && h_carry < dt) ? h_carry : dt;
  double k0[NX];
  M::rhs(mp, aux, x, u, k0);
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
    double hk[S][NX];
#pragma unroll
    for (int n = 0; n < NX; ++n) hk[0][n] = hh * k0[n];
#pragma unroll
    for (int i = 1; i < S; ++i) {
      double xi[NX], k[NX];
#pragma unroll
      for (int n = 0; n < NX; ++n) {
        double acc = x[n];
#pragma unroll
        for (int j = 0; j < i; ++j)
          if (TB::a(i, j) != 0.0) acc = fma(dv.a[i][j], hk[j][n], acc);
        xi[n] = acc;
      }
      M::rhs(mp, aux, xi, u, k);
#pragma unroll
      for (int n = 0; n < NX; ++n) hk[i][n] = hh * k[n];
    }
    double xn[NX], kn[NX];
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      double acc = x[n];
#pragma unroll
      for (int i = 0; i < S; ++i)
        if (TB::b(i) != 0.0) acc = fma(dv.b[i], hk[i][n], acc);
      xn[n] = acc;
    }
    M::rhs(mp, aux, xn, u, kn);                      // FSAL stage: k_{S+1} = f(x_new)
    n_rhs += S;
    double sq = 0.0;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      double err = dv.e[S] * (hh * kn[n]);
#pragma unroll
      for (int i = 0; i < S; ++i)
        if (TB::e(i) != 0.0) err = fma(dv.e[i], hk[i][n], err);
      const double sc = fma(rtol, fmax(fabs(x[n]), fabs(xn[n])), atol);
      const double r = err * rcp_nr(sc);               // sc > 0 and normal whenever atol > 0 or the state is non-zero
      sq = fma(r, r, sq);
    }
    // err = sqrt(sq / NX) is only needed as err <= 1 and err^-1/5 = (sq / NX)^-1/10: no square root
    const double msq = sq * (1.0 / NX);
    const bool finite = isfinite(msq);
    const float grow = finite ? 0.9f * __powf(fmaxf(static_cast<float>(msq), 1e-30f), -0.1f) : 0.2f;
    if (finite && msq <= 1.0) {
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

// ------------------------------------------------------------------ stiff adaptive step, one environment
// Rosenbrock method TB (RODAS4: tableaus.cuh) with the model's analytic Jacobian and its own linear solver
// (M::factor / M::solve), one Jacobian + factorisation per attempted step, six right-hand sides and six solves.
// Same contract as advance_adaptive: per-lane step control on err = RMS(U_6 / (atol + rtol max(|x|,|x'|))), the embedded
// solution is one order lower so h *= clamp(0.9 err^-1/4, 0.2, 6); step carried across intervals; status bits on a
// step-cap or a step that shrinks to nothing.  The stiff counterpart of the reference's CVODES BDF (integrator.py:163-182).
template <class M, class TB>
PCG_DEV void advance_rosenbrock(const typename M::P& mp, double (&x)[M::NX], const double (&u)[M::NU + M::ND], double dt,
                                double rtol, double atol, int max_steps, double& h_carry, int& n_rhs, unsigned& status) {
  constexpr int NX = M::NX, S = TB::S;
  typename M::Aux aux;
  M::prep(mp, u, aux);
  double t = 0.0;
  double h = (h_carry > 0.0 && h_carry < dt) ? h_carry : 0.25 * dt;
  int steps = 0;
  bool rejected = false;
#pragma unroll 1
  while (t < dt) {
    if (steps >= max_steps) { status |= PCG_ST_STEPCAP; break; }
    ++steps;
    const double left = dt - t;
    const bool last = (h >= left * (1.0 - 1e-12));
    const double hh = last ? left : h;
    const double inv_h = 1.0 / hh;
    typename M::Lin w;
    M::factor(mp, aux, x, u, inv_h * (1.0 / TB::gamma), w);
    double U[S][NX], yi[NX];
#pragma unroll
    for (int i = 0; i < S; ++i) {
#pragma unroll
      for (int n = 0; n < NX; ++n) {
        double acc = x[n];
#pragma unroll
        for (int j = 0; j < i; ++j)
          if (TB::a(i, j) != 0.0) acc = fma(TB::a(i, j), U[j][n], acc);
        yi[n] = acc;
      }
      double f[NX];
      M::rhs(mp, aux, yi, u, f);
#pragma unroll
      for (int n = 0; n < NX; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < i; ++j)
          if (TB::c(i, j) != 0.0) acc = fma(TB::c(i, j), U[j][n], acc);
        f[n] = fma(acc, inv_h, f[n]);
      }
      M::solve(w, f);
#pragma unroll
      for (int n = 0; n < NX; ++n) U[i][n] = f[n];
    }
    n_rhs += S + 1;                                  // six right-hand sides and the Jacobian
    double sq = 0.0, xn[NX];
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      xn[n] = yi[n] + U[S - 1][n];                   // yi is the argument of the last stage = the embedded solution
      const double sc = fma(rtol, fmax(fabs(x[n]), fabs(xn[n])), atol);
      const double r = U[S - 1][n] * rcp_nr(sc);
      sq = fma(r, r, sq);
    }
    const double msq = sq * (1.0 / NX);              // err^2
    const bool finite = isfinite(msq);
    const float grow = finite ? 0.9f * __powf(fmaxf(static_cast<float>(msq), 1e-30f), -0.125f) : 0.2f;
    if (finite && msq <= 1.0) {
      t = last ? dt : t + hh;
#pragma unroll
      for (int n = 0; n < NX; ++n) x[n] = xn[n];
      const double hn = hh * static_cast<double>(fminf(rejected ? 1.0f : 6.0f, grow));
      h = last ? fmax(h, hn) : hn;
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
template <class M, class TB, bool ADAPTIVE, bool GENERAL, int EPT>
__global__ void __launch_bounds__(kBlock, (EPT == 1 && !ADAPTIVE) ? M::MIN_BLOCKS_EPT1 : M::MIN_BLOCKS)
rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io,
               const __grid_constant__ Derived<M, TB> dv, const int64_t n, const int t0, const int k_steps) {
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND;
  constexpr int NSPM = GENERAL ? PCG_MAX_NSP : 2;              // setpoints handled by this instantiation
  constexpr int NDOM = GENERAL ? PCG_MAX_ND : 1;
  constexpr int NTLM = GENERAL ? PCG_MAX_NTAIL : 1;
  constexpr int NCONM = GENERAL ? PCG_MAX_NCON : 4;            // constraints handled by this instantiation
  __shared__ __align__(16) double icdf_s[kIcdfDoubles];        // inverse-normal-CDF table (philox.cuh)
  if (d.noise == PCG_NOISE_PHILOX) {                           // uniform: the whole CTA takes the same path
    fill_icdf_tab(icdf_s);
    __syncthreads();
  }
  const int64_t env0 = (static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x) * EPT;
  if (env0 >= n) return;                                       // EPT == 2 implies even n: both envs valid

  const int flags = d.flags;
  const bool norm_a = flags & PCG_F_NORMALISE_A;
  const bool norm_o = flags & PCG_F_NORMALISE_O;
  const bool a_delta = norm_a && (flags & PCG_F_A_DELTA);                      // pcgym.py:376
  const bool penal = flags & PCG_F_R_PENALTY;
  const bool done_on_v = flags & PCG_F_DONE_ON_VIOLATION;
  const bool dist_on = GENERAL && d.nd > 0;
  const int nsp = d.nsp, n_con = d.n_con;
  const int nd_obs = GENERAL ? d.nd_obs : 0, n_tail = GENERAL ? d.n_tail : 0;
  const int nsp_obs = (flags & PCG_F_SP_HIDDEN) ? 0 : nsp;                     // setpoint slots of the observation
  const int n_obs = NX + nsp_obs + nd_obs + n_tail;
  const int noise = d.noise;

  // ---- carried state
  double x[EPT][NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v[EPT];
    ld_row<EPT>(io.x + i * n + env0, v);
#pragma unroll
    for (int e = 0; e < EPT; ++e) x[e][i] = v[e];
  }
  double a_save[EPT][NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    double v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = 0.0;
    if (a_delta) ld_row<EPT>(io.a_save + j * n + env0, v);
#pragma unroll
    for (int e = 0; e < EPT; ++e) a_save[e][j] = v[e];
  }
  double tail[EPT][NTLM];
#pragma unroll
  for (int j = 0; j < NTLM; ++j) {
    double v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = 0.0;
    if (GENERAL && j < n_tail) ld_row<EPT>(io.tail + j * n + env0, v);
#pragma unroll
    for (int e = 0; e < EPT; ++e) tail[e][j] = v[e];
  }
  bool done[EPT];
  unsigned status[EPT];
  int n_rhs[EPT];
  double h_carry[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    done[e] = (io.done_state && !(flags & PCG_F_EPISODE_START)) ? (io.done_state[env0 + e] != 0) : false;
    status[e] = 0;
    n_rhs[e] = 0;
    h_carry[e] = (ADAPTIVE && io.h_carry) ? io.h_carry[env0 + e] : 0.0;
  }

  // ---- model constants: the descriptor's (constant bank) or, GENERAL only, built per env from io.theta
  constexpr int NPL = GENERAL ? EPT : 1;
  typename M::P p_env[NPL];
  double d_def[NPL][ND > 0 ? ND : 1];          // what an unsupplied disturbance channel falls back to
  const bool per_env = GENERAL && io.theta != nullptr;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    p_env[e] = dv.p;
#pragma unroll
    for (int i = 0; i < ND; ++i) d_def[e][i] = d.d_default[i];
    if (per_env) {
      double th[M::NP];
#pragma unroll
      for (int i = 0; i < M::NP; ++i) th[i] = io.theta[static_cast<int64_t>(i) * n + env0 + e];
      p_env[e] = M::make(th);
#pragma unroll
      for (int i = 0; i < ND; ++i)
        if (M::dpar(i) >= 0) d_def[e][i] = th[M::dpar(i)];
    }
  }
  const bool ocp = d.reward == PCG_REWARD_OCP;
  double u_prev[EPT][NU];
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    double v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = 0.0;
    if (ocp && io.u_prev) ld_row<EPT>(io.u_prev + j * n + env0, v);
#pragma unroll
    for (int e = 0; e < EPT; ++e) u_prev[e][j] = v[e];
  }
  bool u_prev_valid = (d.ocp_mode & 3) != 2;

  // ---- per-thread stream pointers, advanced by a uniform stride every interval
  const double* pa = io.a + env0;
  const double* pz = io.z ? io.z + env0 : nullptr;
  double* px = io.x_traj ? io.x_traj + env0 : nullptr;
  double* po = io.obs ? io.obs + env0 : nullptr;
  double* pr = io.reward ? io.reward + env0 : nullptr;
  double* pg = io.g ? io.g + env0 : nullptr;
  uint8_t* pd = io.done ? io.done + env0 : nullptr;
  uint8_t* pv = io.violated ? io.violated + env0 : nullptr;

  // setpoints: SP[t] is shown in the observation, SP[t+1] enters the reward (pcgym.py:432-441, 555).
  // psp[j] walks the schedule one time index per interval; two values are carried in registers.
  const double* psp[NSPM];
  double sp_cur[NSPM][EPT], sp_nxt[NSPM][EPT];
#pragma unroll
  for (int j = 0; j < NSPM; ++j) {
    psp[j] = nullptr;
#pragma unroll
    for (int e = 0; e < EPT; ++e) sp_cur[j][e] = sp_nxt[j][e] = 0.0;
    if (j < nsp) {
      psp[j] = io.sp + j * io.sp_sj + env0 * io.sp_se + static_cast<int64_t>(t0) * io.sp_st;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        sp_cur[j][e] = psp[j][e * io.sp_se];
        sp_nxt[j][e] = psp[j][e * io.sp_se + io.sp_st];
      }
      psp[j] += 2 * io.sp_st;                                 // -> SP[t0 + 2]
    }
  }

  // register double buffer for the streamed actions
  double a_nxt[NU][EPT];
#pragma unroll
  for (int j = 0; j < NU; ++j) ld_row<EPT>(pa + j * n, a_nxt[j]);
  const PhiloxKeys& key = dv.rk;

#pragma unroll 1
  for (int k = 0; k < k_steps; ++k) {
    const int t = t0 + k;
    const int t_new = t + 1;                                                   // pcgym.py:441
    double a_raw[NU][EPT];
#pragma unroll
    for (int j = 0; j < NU; ++j)
#pragma unroll
      for (int e = 0; e < EPT; ++e) a_raw[j][e] = a_nxt[j][e];
    pa += static_cast<int64_t>(NU) * n;
    if (k + 1 < k_steps) {
#pragma unroll
      for (int j = 0; j < NU; ++j) ld_row<EPT>(pa + j * n, a_nxt[j]);
    }
    double sp_fut[NSPM][EPT];                                                  // SP[t+2], consumed next interval
#pragma unroll
    for (int j = 0; j < NSPM; ++j) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) sp_fut[j][e] = sp_nxt[j][e];
      if (j < nsp && t + 2 < d.N) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) sp_fut[j][e] = psp[j][e * io.sp_se];
        psp[j] += io.sp_st;
      }
    }

    // ---- action -> physical input (pcgym.py:372-383)
    double uk[EPT][NUK];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double a = a_raw[j][e];
        if (norm_a) a = fma(a, d.a_scale[j], d.a_off[j]);
        if (a_delta) {
          a = fma(a, d.a_scale[j], d.a_off[j]);     // normalised a second time, as the reference does
          a = a_save[e][j] + a;                     // the UNclipped sum is integrated ...
          a_save[e][j] = fmin(fmax(a, d.a_act_lo[j]), d.a_act_hi[j]);   // ... only a_save is clipped
        }
        uk[e][j] = a;
      }
      // ---- disturbances at t+1 (pcgym.py:386-412); defaults when inactive so the RHS needs no branch
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        double v = GENERAL ? d_def[e][i] : d.d_default[i];
        if (dist_on && d.d_src[i] >= 0)
          v = io.dist[static_cast<int64_t>(t + 1) * io.d_st + d.d_src[i] * io.d_sj + (env0 + e) * io.d_se];
        uk[e][NU + i] = v;
      }
    }
    double dslot[EPT][NDOM];                          // observed disturbance slots of the state vector
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int j = 0; j < NDOM; ++j) {
        double v = 0.0;
        if (GENERAL && ND > 0 && j < nd_obs) {
#pragma unroll
          for (int i = 0; i < ND; ++i) v = (d.d_obs[j] == i) ? uk[e][NU + i] : v;
        }
        dslot[e][j] = v;
      }

    // ---- constraint evaluator over state' = [x | sp | dslot | tail] and uk; g > 0 = violated.
    // Fully unrolled over the constraint index so G and h are loaded from fixed constant-bank offsets.
    auto constraints = [&](const double (&sps)[NSPM][EPT], double* gout, bool (&viol)[EPT]) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) viol[e] = false;
#pragma unroll
      for (int c = 0; c < NCONM; ++c) {
        if (c < n_con) {
          double g[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) g[e] = d.con_h[c];
          // dense over the model states and inputs (zero coefficients cost one DFMA but no predicate
          // logic); the optional observation slots are sparse and stay masked
#pragma unroll
          for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int e = 0; e < EPT; ++e) g[e] = fma(d.con_gx[c][i], x[e][i], g[e]);
#pragma unroll
          for (int i = 0; i < NUK; ++i)
#pragma unroll
            for (int e = 0; e < EPT; ++e) g[e] = fma(d.con_gu[c][i], uk[e][i], g[e]);
          if constexpr (GENERAL) {
            const uint64_t mx = d.con_mask_x[c];
            if (mx >> NX) {                           // rare: constraint reads SP / disturbance / tail slots
#pragma unroll
              for (int e = 0; e < EPT; ++e) {
#pragma unroll
                for (int j = 0; j < NSPM; ++j)
                  if (j < nsp && (mx >> (NX + j) & 1u)) g[e] = fma(d.con_gx[c][NX + j], sps[j][e], g[e]);
#pragma unroll
                for (int j = 0; j < NDOM; ++j)
                  if (j < nd_obs && (mx >> (NX + nsp + j) & 1u)) g[e] = fma(d.con_gx[c][NX + nsp + j], dslot[e][j], g[e]);
#pragma unroll
                for (int j = 0; j < NTLM; ++j)
                  if (j < n_tail && (mx >> (NX + nsp + nd_obs + j) & 1u))
                    g[e] = fma(d.con_gx[c][NX + nsp + nd_obs + j], tail[e][j], g[e]);
              }
            }
          }
#pragma unroll
          for (int e = 0; e < EPT; ++e) viol[e] = viol[e] || (g[e] > 0.0);   // NaN compares false, as np.any(g > 0)
          if (gout) st_row<EPT>(gout + c * n, g);
        }
      }
    };

    // ---- t == 0 pre-check (pcgym.py:416-420): records cons_info[:,0,:], may set done
    if (t == 0 && n_con > 0) {
      double sp0[NSPM][EPT];
#pragma unroll
      for (int j = 0; j < NSPM; ++j)
#pragma unroll
        for (int e = 0; e < EPT; ++e) sp0[j][e] = d.sp_init[j];
      bool v0[EPT];
      constraints(sp0, io.g0 ? io.g0 + env0 : nullptr, v0);
#pragma unroll
      for (int e = 0; e < EPT; ++e)
        if (v0[e] && done_on_v) done[e] = true;
    }

    // ---- Gaussian draws for this interval.  They depend on (seed, env, t) only, so they are issued in the
    // same basic block as the integrator stages: independent work that fills the FP64 dependency stalls.
    double z[EPT][NX];
    auto draw = [&]() {
#pragma unroll
      for (int e = 0; e < EPT; ++e)
#pragma unroll
        for (int i = 0; i < NX; ++i) z[e][i] = 0.0;
      if (noise == PCG_NOISE_PHILOX || noise == PCG_NOISE_PHILOX_F32) {
        const uint64_t gid0 = static_cast<uint64_t>(d.env_id0 + env0);
#pragma unroll
        for (int i = 0; i < NX; i += 2) {
          double za[EPT], zb[EPT];
          if (noise == PCG_NOISE_PHILOX) draw_state_pair<EPT, false>(gid0, static_cast<uint32_t>(t_new), i / 2, key, icdf_s, za, zb);
          else draw_state_pair<EPT, true>(gid0, static_cast<uint32_t>(t_new), i / 2, key, icdf_s, za, zb);
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            z[e][i] = za[e];
            if (i + 1 < NX) z[e][i + 1] = zb[e];
          }
        }
      } else if (noise == PCG_NOISE_EXTERNAL) {
#pragma unroll
        for (int i = 0; i < NX; ++i)
          if (d.noise_pct[i] != 0.0) {
            double v[EPT];
            ld_row<EPT>(pz + i * n, v);
#pragma unroll
            for (int e = 0; e < EPT; ++e) z[e][i] = v[e];
          }
      }
    };

    // ---- integrate over the control interval, input held (integrator.py:90-107)
    if constexpr (ADAPTIVE && TB::EMBEDDED) {
      draw();
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        if constexpr (is_rosenbrock<TB>::value)
          advance_rosenbrock<M, TB>(GENERAL ? p_env[e % NPL] : dv.p, x[e], uk[e], d.dt, d.rtol, d.atol, d.max_steps,
                                    h_carry[e], n_rhs[e], status[e]);
        else
          advance_adaptive<M, TB>(dv, GENERAL ? p_env[e % NPL] : dv.p, x[e], uk[e], d.dt, d.rtol, d.atol, d.max_steps,
                                  h_carry[e], n_rhs[e], status[e]);
      }
    } else {
      typename M::Aux aux[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) M::prep(GENERAL ? p_env[e % NPL] : dv.p, uk[e], aux[e]);
      draw();
      if constexpr (GENERAL) {
#pragma unroll 1
        for (int s = 0; s < d.n_sub; ++s) rk_step_fixed<M, TB, EPT>(dv, p_env, aux, x, uk);
      } else if (d.n_sub == 1) {
        rk_step_fixed<M, TB, EPT>(dv, aux, x, uk);
      } else {
#pragma unroll 1
        for (int s = 0; s < d.n_sub; ++s) rk_step_fixed<M, TB, EPT>(dv, aux, x, uk);
      }
    }

    // ---- constraints on the new state (pcgym.py:444-446), done (pcgym.py:448-449)
    bool violated[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) violated[e] = false;
    if (n_con > 0) {
      constraints(sp_cur, pg, violated);
#pragma unroll
      for (int e = 0; e < EPT; ++e)
        if (violated[e] && done_on_v) done[e] = true;
    }
    if (t_new == d.N - 1) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) done[e] = true;
    }

    // ---- reward on the noise-free state (pcgym.py:470-482)
    double r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) r[e] = 0.0;
    if (d.reward == PCG_REWARD_SP) {                                           // pcgym.py:535-558
#pragma unroll
      for (int j = 0; j < NSPM; ++j)
        if (j < nsp) {
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const double err = pick<NX>(x[e], d.sp_state[j]) - sp_nxt[j][e];
            r[e] = fma(-(err * err), d.r_scale[j], r[e]);
            if (penal && violated[e]) r[e] -= d.penalty;                       // once per SP key
          }
        }
    } else if (ocp) {
      // fused "OCP tracking" custom reward of the paper scripts (custom_reward.py:3-39): evaluated on the
      // NOISY physical observation, setpoint SP[t+1], control cost on the normalised physical input
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        double cost = 0.0;
#pragma unroll
        for (int j = 0; j < NSPM; ++j)
          if (j < nsp) {
            const int si = d.sp_state[j];
            double xo = 0.0;
            if (NX >= 3 && (d.ocp_mode & PCG_OCP_MOMENTS)) {
              // crystallization script (cryst_train.py:17-48): the tracked CV and Ln are DERIVED from the noisy
              // moments, CV = sqrt(mu2 mu0 / mu1^2 - 1) against the first setpoint, Ln = mu1 / mu0 against the second
              double o[3];
#pragma unroll
              for (int i = 0; i < 3; ++i) o[i] = fma(z[e][i % NX] * x[e][i % NX], d.noise_pct[i % NX], x[e][i % NX]);
              xo = (j == 0) ? sqrt_fast(o[2] * o[0] * rcp_nr(o[1] * o[1]) - 1.0) : o[1] * rcp_nr(o[0]);
            } else {
#pragma unroll
              for (int i = 0; i < NX; ++i) {
                const double ox = fma(z[e][i] * x[e][i], d.noise_pct[i], x[e][i]);
                xo = (si == i) ? ox : xo;
              }
            }
            const double en = (xo - sp_nxt[j][e]) * d.ocp_inv_span[j];
            cost = fma(en * en, d.r_scale[j], cost);
          }
#pragma unroll
        for (int j = 0; j < NU; ++j) {
          const double inv = 0.5 * rcp_nr(d.a_scale[j]);                       // 1 / (a_hi - a_lo)
          const double un = (uk[e][j] - (d.a_off[j] - d.a_scale[j])) * inv;
          const double up = ((u_prev_valid ? u_prev[e][j] : uk[e][j]) - (d.a_off[j] - d.a_scale[j])) * inv;
          const double du = (d.ocp_mode & 3) == 0 ? un : un - up;
          cost = fma(d.ocp_R * du, du, cost);
          u_prev[e][j] = uk[e][j];
        }
        r[e] = -cost;
      }
      u_prev_valid = true;
    } else if (d.reward == PCG_REWARD_BATCH) {                                 // pcgym.py:502-532
      if (t_new == d.N - 1) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
#pragma unroll
          for (int i = 0; i < NX; ++i) r[e] = fma(d.batch_w[i], x[e][i], r[e]);
          if (penal && violated[e]) r[e] -= d.penalty;
        }
      }
    }

    // ---- outputs (evict-first: never re-read by this kernel)
    if (px) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) v[e] = x[e][i];
        st_row<EPT>(px + i * n, v);
      }
      px += static_cast<int64_t>(NX) * n;
    }
    if (po) {
      const uint32_t zmask = d.obs_zero_mask;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          // measurement noise on the model states only: obs = x + z x pct (pcgym.py:453-466)
          const double ox = fma(z[e][i] * x[e][i], d.noise_pct[i], x[e][i]);
          v[e] = norm_o ? fma(ox, d.o_scale[i], d.o_off[i]) : ox;             // pcgym.py:483-493
          if (zmask >> i & 1u) v[e] = 0.0;                                     // pcgym.py:495-498
        }
        st_row<EPT>(po + i * n, v);
      }
#pragma unroll
      for (int j = 0; j < NSPM; ++j)
        if (j < nsp_obs) {
          const int i = NX + j;
          double v[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) v[e] = norm_o ? fma(sp_cur[j][e], d.o_scale[i], d.o_off[i]) : sp_cur[j][e];
          st_row<EPT>(po + i * n, v);
        }
      if (GENERAL) {
#pragma unroll
        for (int j = 0; j < NDOM; ++j)
          if (j < nd_obs) {
            const int i = NX + nsp_obs + j;
            double v[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) v[e] = norm_o ? fma(dslot[e][j], d.o_scale[i], d.o_off[i]) : dslot[e][j];
            st_row<EPT>(po + i * n, v);
          }
#pragma unroll
        for (int j = 0; j < NTLM; ++j)
          if (j < n_tail) {
            const int i = NX + nsp_obs + nd_obs + j;
            double v[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) v[e] = norm_o ? fma(tail[e][j], d.o_scale[i], d.o_off[i]) : tail[e][j];
            st_row<EPT>(po + i * n, v);
          }
      }
      po += static_cast<int64_t>(n_obs) * n;
    }
    if (pr) { st_row<EPT>(pr, r); pr += n; }
    if (pd) { st_flags<EPT>(pd, done); pd += n; }
    if (pv) { st_flags<EPT>(pv, violated); pv += n; }
    if (pg) pg += static_cast<int64_t>(n_con) * n;
    if (pz) pz += static_cast<int64_t>(NX) * n;
#pragma unroll
    for (int j = 0; j < NSPM; ++j)
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        sp_cur[j][e] = sp_nxt[j][e];
        sp_nxt[j][e] = sp_fut[j][e];
      }
  }

  // ---- carried state back to memory; non-finite states are flagged once per launch (NaN is sticky)
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      v[e] = x[e][i];
      if (!isfinite(v[e])) status[e] |= PCG_ST_NONFINITE;
    }
    st_row<EPT>(io.x + i * n + env0, v);
  }
  if (a_delta) {
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double v[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) v[e] = a_save[e][j];
      st_row<EPT>(io.a_save + j * n + env0, v);
    }
  }
  if (ocp && io.u_prev) {
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double v[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) v[e] = u_prev[e][j];
      st_row<EPT>(io.u_prev + j * n + env0, v);
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    if (io.done_state) io.done_state[env0 + e] = done[e] ? 1 : 0;
    if (io.status && status[e]) io.status[env0 + e] |= static_cast<uint8_t>(status[e]);
    if (ADAPTIVE && io.h_carry) io.h_carry[env0 + e] = h_carry[e];
    if (io.n_rhs) io.n_rhs[env0 + e] += ADAPTIVE ? n_rhs[e] : k_steps * d.n_sub * TB::S;
  }
}

}  // namespace pcg
