// Warp-specialised form of the straight-line control-interval kernel (same eligibility as rollout_fast.cuh).
//
// Why it exists (profiles/README.md): the monolithic kernel asks ptxas to interleave, statically and inside one
// basic block, three phases that have nothing to do with each other -- the Gaussian draw of interval k, the
// integrator stages of interval k and the bookkeeping of interval k-1 -- and pays 255 registers for it; at
// 65 536 envs each SM sub-partition then holds ~2 warps and the FP64 pipe idles half of the time waiting on
// fixed-latency dependencies.  Here the phases live in DIFFERENT warps of one CTA and the hardware scheduler
// interleaves them dynamically:
//
//   integrator warp  (serial in time)   : wait action -> un-normalise -> M::prep -> RK stages -> x_end to smem
//   finisher warp    (parallel in time) : global loads of the action rows (prefetched kWsRing intervals ahead)
//                                         -> smem; Philox + Box-Muller; constraints, done, reward, noisy
//                                         normalised observation; every global store
//
// The two warps of a pair own the same 64 environments (lane L: envs 2L, 2L+1, so rows stay 128-bit in HBM and
// in shared memory) and hand tiles through a ring of kWsRing slots guarded by mbarriers:
//   full_a[s]  finisher -> integrator : action rows of interval k are in slot s (and x[s] may be overwritten)
//   full_x[s]  integrator -> finisher : end state of interval k is in slot s (and a[s] has been consumed)
// so the integrator never waits for the bookkeeping unless it runs kWsRing intervals ahead.
#pragma once
#include "rollout_fast.cuh"

namespace pcg {

constexpr int kWsRing = 8;         // power of two

PCG_DEV uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
PCG_DEV void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
PCG_DEV void mbar_arrive(uint64_t* b) {          // release.cta: this lane's earlier smem writes are visible to waiters
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
PCG_DEV bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(smem_u32(b)), "r"(parity)
               : "memory");
  return ok != 0;
}
PCG_DEV void mbar_wait(uint64_t* b, uint32_t parity) {   // acquire.cta; retries spin (pair-per-CTA kernel: waits are rare)
  while (!mbar_try(b, parity)) {}
}
// The same wait for the per-SM kernel, where one of the two roles is ahead of the other most of the time: measured
// (ncu source page, profiles/README.md r2) try_wait comes back within ~7 cycles, and the retry loops of the waiting
// role issued 34 M of the kernel's 87 M warp instructions -- issue slots taken from the role that sets the pace.
// So a failed first try backs off: PCG_WSM_WAIT 1 = nanosleep between tries (the warp leaves the scheduler),
// 2 = try_wait with a suspend-time hint, 0 = spin (A/B).
#ifndef PCG_WSM_WAIT
#define PCG_WSM_WAIT 1
#endif
#ifndef PCG_WSM_SLEEP_NS
#define PCG_WSM_SLEEP_NS 64
#endif
PCG_DEV void mbar_wait_backoff(uint64_t* b, uint32_t parity) {
  if (mbar_try(b, parity)) return;
#if PCG_WSM_WAIT == 1
  do { __nanosleep(PCG_WSM_SLEEP_NS); } while (!mbar_try(b, parity));
#elif PCG_WSM_WAIT == 2
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity), "r"(0x989680)
                 : "memory");
  } while (!ok);
#else
  while (!mbar_try(b, parity)) {}
#endif
}

// models whose integrator-warp RHS reads the shared-memory 2^(j/256) table (exp_tab_split): M::EXP_TAB
template <class M, class = void>
struct uses_exp_tab : std::false_type {};
template <class M>
struct uses_exp_tab<M, std::enable_if_t<M::EXP_TAB>> : std::true_type {};

// models whose table is indexed directly over their whole argument range (exp_wide_split): M::EXP_WIDE entries
template <class M, class = void>
struct has_wide_tab : std::false_type {};
template <class M>
struct has_wide_tab<M, std::enable_if_t<(M::EXP_WIDE > 0)>> : std::true_type {};
template <class M>
constexpr int exp_tab_doubles() {
  if constexpr (has_wide_tab<M>::value) return M::EXP_WIDE;
  else return uses_exp_tab<M>::value ? 256 : 1;
}

template <class M>
PCG_DEV void fill_exp_tab(double* tab, const typename M::P& p) {   // whole CTA, before the __syncthreads that publishes the barriers
  if constexpr (has_wide_tab<M>::value) {
    M::fill_tab(tab, p);
  } else if constexpr (uses_exp_tab<M>::value) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = kExp2Tab[i];
  }
}
// PCG_ST_RANGE for a state outside the band a wide-table model checks at interval ends; 0 for every other model
template <class M>
PCG_DEV unsigned range_flag(const typename M::P& p, const double* x) {
  if constexpr (has_wide_tab<M>::value) return M::range_flag(p, x);
  else return 0u;
}
template <class M>
PCG_DEV void bind_exp_tab(typename M::Aux& a, const double* tab) {
  if constexpr (uses_exp_tab<M>::value) a.tab = tab;
}

template <class M, int RING = kWsRing>
struct WsPair {
  double2 a[RING][M::NU][32];
  double2 x[RING][M::NX][32];
  uint64_t full_a[RING];
  uint64_t full_x[RING];
};

// NSP = number of setpoints (1 or 2) at compile time: no dead reward / setpoint work (see rollout_wsm.cuh)
template <class M, class TB, int NOISE, int W, int NSP>
__global__ void __launch_bounds__(64 * W, W == 1 ? 7 : 3)
ws_rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io,
                  const __grid_constant__ Derived<M, TB> dv, const int64_t n, const int t0, const int k_steps) {
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND, EPT = 2, R = kWsRing;
  __shared__ __align__(16) WsPair<M> sm[W];
  __shared__ double exp_tab[exp_tab_doubles<M>()];
  __shared__ __align__(16) double icdf_s[NOISE == PCG_NOISE_PHILOX ? kIcdfDoubles : 2];
  fill_exp_tab<M>(exp_tab, dv.p);
  if constexpr (NOISE == PCG_NOISE_PHILOX) fill_icdf_tab(icdf_s);
  // Role placement.  A warp's scheduler (SM sub-partition) is its hardware slot %warpid mod 4 and the CTAs of an
  // SM occupy consecutive slots, so a fixed (integrator, finisher) order would put every integrator warp -- the
  // FP64-heavy role -- on the same two sub-partitions.  Warp 2p / 2p+1 form pair p; which of the two integrates
  // alternates with the slot pair so that after four pairs every sub-partition holds one warp of each role.
  // (The flip is decided once per CTA from warp 0's slot and shared, so both warps of a pair always agree.)
  __shared__ int flip_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    unsigned slot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    const unsigned k = slot >> 1;
    flip_s = static_cast<int>((k ^ (k >> 1)) & 1u);
  }
  const int pair = warp >> 1;
  WsPair<M>& sp = sm[pair];
  const int64_t pair_base = (static_cast<int64_t>(blockIdx.x) * W + pair) * 64;
  const int64_t left = (n - pair_base) / 2;                    // vector path: n is even
  const int n_act = left < 0 ? 0 : (left > 32 ? 32 : static_cast<int>(left));
  if ((warp & 1) == 0 && lane == 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mbar_init(&sp.full_a[s], n_act > 0 ? n_act : 1);
      mbar_init(&sp.full_x[s], n_act > 0 ? n_act : 1);
    }
  }
  __syncthreads();
  const bool integ = ((warp & 1) ^ flip_s ^ (pair & 1)) == 0;
  if (lane >= n_act) return;
  const int64_t env0 = pair_base + 2 * lane;

  auto to_uk = [&](const double (&a_raw)[NU][EPT], double (&uk)[EPT][NUK]) {   // pcgym.py:372-375, 386-414
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int j = 0; j < NU; ++j) uk[e][j] = fma(a_raw[j][e], d.a_scale[j], d.a_off[j]);
#pragma unroll
      for (int i = 0; i < ND; ++i) uk[e][NU + i] = d.d_default[i];
    }
  };

  if (integ) {
    // ================================================================ integrator warp
    double x[EPT][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v[EPT];
      ld_row<EPT>(io.x + i * n + env0, v);
#pragma unroll
      for (int e = 0; e < EPT; ++e) x[e][i] = v[e];
    }
#pragma unroll 1
    for (int k = 0; k < k_steps; ++k) {
      const int s = k & (R - 1);
      mbar_wait(&sp.full_a[s], (k / R) & 1);
      double a_raw[NU][EPT];
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        const double2 t = sp.a[s][j][lane];
        a_raw[j][0] = t.x;
        a_raw[j][1] = t.y;
      }
      double uk[EPT][NUK];
      to_uk(a_raw, uk);
      typename M::Aux aux[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        M::prep(dv.p, uk[e], aux[e]);
        bind_exp_tab<M>(aux[e], exp_tab);
      }
      rk_step_fixed<M, TB, EPT>(dv, aux, x, uk);                               // integrator.py:90-107
#pragma unroll
      for (int i = 0; i < NX; ++i) sp.x[s][i][lane] = make_double2(x[0][i], x[1][i]);
      mbar_arrive(&sp.full_x[s]);
    }
    unsigned status[EPT] = {0, 0};
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
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      if (io.status && status[e]) io.status[env0 + e] |= static_cast<uint8_t>(status[e]);
      if (io.n_rhs) io.n_rhs[env0 + e] += k_steps * TB::S;
    }
    return;
  }

  // ================================================================== finisher warp
  const int flags = d.flags;
  const bool done_on_v = flags & PCG_F_DONE_ON_VIOLATION;
  const int nsp = d.nsp, n_con = d.n_con;
  const int nsp_obs = (flags & PCG_F_SP_HIDDEN) ? 0 : nsp, n_obs = NX + nsp_obs;
  const double pen = d.penalty;
  const PhiloxKeys& key = dv.rk;
  const int t_last = d.N - 1;

  bool done[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) done[e] = (io.done_state && !(d.flags & PCG_F_EPISODE_START)) ? (io.done_state[env0 + e] != 0) : false;

  double* px = io.x_traj + env0;
  double* po = io.obs + env0;
  double* pr = io.reward + env0;
  double* pg = io.g ? io.g + env0 : nullptr;
  uint8_t* pd = io.done + env0;
  uint8_t* pv = io.violated + env0;

  // g = G [x; uk] + h, dense; rows >= n_con are zero-filled by the host so they can never flag
  bool con_u = false;
#pragma unroll
  for (int c = 0; c < kFastNcon; ++c) con_u = con_u || d.con_mask_u[c] != 0;
  auto constraints = [&](const double (&xs)[EPT][NX], const double (&uk)[EPT][NUK], double* gout, bool (&viol)[EPT]) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) viol[e] = false;
#pragma unroll
    for (int c = 0; c < kFastNcon; ++c) {
      double g[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        double acc = d.con_h[c];
#pragma unroll
        for (int i = 0; i < NX; ++i) acc = fma(d.con_gx[c][i], xs[e][i], acc);
        if (con_u) {                                     // uniform: most constraints read states only
#pragma unroll
          for (int i = 0; i < NUK; ++i) acc = fma(d.con_gu[c][i], uk[e][i], acc);
        }
        g[e] = acc;
        viol[e] = viol[e] || (acc > 0.0);
      }
      if (gout != nullptr && c < n_con) st_row<EPT>(gout + c * n, g);
    }
  };

  // initial state for the t == 0 pre-check, read BEFORE the integrator can be released (it overwrites io.x at the end)
  const bool pre_check = t0 == 0 && n_con > 0;
  double x0[EPT][NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v[EPT] = {0.0, 0.0};
    if (pre_check) ld_row<EPT>(io.x + i * n + env0, v);
    x0[0][i] = v[0];
    x0[1][i] = v[1];
  }

  // ---- fill the ring with the first intervals' action rows
  const double* pa = io.a + env0;
  double a0[NU][EPT];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    if (s < k_steps) {
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        double v[EPT];
        ld_row<EPT>(pa + (static_cast<int64_t>(s) * NU + j) * n, v);
        sp.a[s][j][lane] = make_double2(v[0], v[1]);
        if (s == 0) { a0[j][0] = v[0]; a0[j][1] = v[1]; }
      }
      mbar_arrive(&sp.full_a[s]);
    }
  }
  pa += static_cast<int64_t>(R < k_steps ? R : k_steps - 1) * NU * n;           // row of interval min(R, K-1)

  // ---- t == 0 pre-check (pcgym.py:416-420) on the initial state with the first interval's input
  if (pre_check) {
    double uk0[EPT][NUK];
    to_uk(a0, uk0);
    bool v0[EPT];
    constraints(x0, uk0, io.g0 ? io.g0 + env0 : nullptr, v0);
#pragma unroll
    for (int e = 0; e < EPT; ++e) done[e] = done[e] || (v0[e] && done_on_v);
  }

  // setpoint schedule pointers (rows >= nsp alias row 0: loaded, never used)
  const double* psp[NSP];
  double sp_cur[NSP][EPT], sp_nxt[NSP][EPT];
#pragma unroll
  for (int j = 0; j < NSP; ++j) {
    psp[j] = io.sp + j * io.sp_sj + env0 * io.sp_se + static_cast<int64_t>(t0) * io.sp_st;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      sp_cur[j][e] = psp[j][e * io.sp_se];
      sp_nxt[j][e] = psp[j][e * io.sp_se + io.sp_st];
    }
    psp[j] += io.sp_st;                                       // -> SP[t0 + 1]; advanced before each use
  }

#pragma unroll 1
  for (int k = 0; k < k_steps; ++k) {
    const int s = k & (R - 1);
    const int t = t0 + k, t_new = t + 1;                                       // pcgym.py:441
    // action rows of interval k + R (pointer clamped at the last interval instead of branching)
    const bool more = k + R < k_steps;
    double a_pre[NU][EPT];
#pragma unroll
    for (int j = 0; j < NU; ++j) ld_row<EPT>(pa + j * n, a_pre[j]);
    pa += (k + R + 1 < k_steps) ? static_cast<int64_t>(NU) * n : 0;

    double spc[NSP][EPT], spn[NSP][EPT];           // SP[t] (observation), SP[t+1] (reward)
#pragma unroll
    for (int j = 0; j < NSP; ++j) {
      psp[j] += (t + 2 < d.N) ? io.sp_st : 0;                                  // -> SP[min(t + 2, N - 1)]
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        spc[j][e] = sp_cur[j][e];
        spn[j][e] = sp_nxt[j][e];
        sp_cur[j][e] = sp_nxt[j][e];
        sp_nxt[j][e] = psp[j][e * io.sp_se];
      }
    }

    // Gaussian draws of this interval: a pure function of (seed, env, t), computed while the integrator works
    double z[EPT][NX];
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int i = 0; i < NX; ++i) z[e][i] = 0.0;
    if constexpr (NOISE != PCG_NOISE_OFF) {
      const uint64_t gid0 = static_cast<uint64_t>(d.env_id0 + env0);           // even: launch.cuh checks env_id0
#pragma unroll
      for (int i = 0; i < NX; i += 2) {
        double za[EPT], zb[EPT];
        draw_state_pair<EPT, NOISE == PCG_NOISE_PHILOX_F32>(gid0, static_cast<uint32_t>(t_new), i / 2, key, icdf_s, za, zb);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          z[e][i] = za[e];
          if (i + 1 < NX) z[e][i + 1] = zb[e];
        }
      }
    }

    mbar_wait(&sp.full_x[s], (k / R) & 1);
    double xe[EPT][NX], a_raw[NU][EPT], uk[EPT][NUK];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const double2 v = sp.x[s][i][lane];
      xe[0][i] = v.x;
      xe[1][i] = v.y;
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const double2 v = sp.a[s][j][lane];
      a_raw[j][0] = v.x;
      a_raw[j][1] = v.y;
    }
    to_uk(a_raw, uk);

    bool violated[EPT];
    constraints(xe, uk, pg, violated);                                         // pcgym.py:444-446
    const bool last = t_new == t_last;                                         // pcgym.py:448-449
    double r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      done[e] = done[e] || (violated[e] && done_on_v) || last;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NSP; ++j) {
        const double err = pick<NX>(xe[e], d.sp_state[j]) - spn[j][e];
        acc = fma(-(err * err), d.r_scale[j], acc);                            // pcgym.py:555
        acc -= violated[e] ? pen : 0.0;                                        // once per SP key (pcgym.py:556)
      }
      r[e] = acc;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v[EPT], w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        v[e] = xe[e][i];
        double ox = (NOISE != PCG_NOISE_OFF) ? fma(z[e][i] * xe[e][i], d.noise_pct[i], xe[e][i]) : xe[e][i];
        w[e] = fma(ox, d.o_scale[i], d.o_off[i]);          // pcgym.py:483-493, 495-498 (folded: fold_desc)
      }
      st_row<EPT>(px + i * n, v);
      st_row<EPT>(po + i * n, w);
    }
#pragma unroll
    for (int j = 0; j < NSP; ++j) {
      double w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) w[e] = fma(spc[j][e], d.o_scale[NX + j], d.o_off[NX + j]);
      if (j < nsp_obs) st_row<EPT>(po + (NX + j) * n, w);
    }
    st_row<EPT>(pr, r);
    st_flags<EPT>(pd, done);
    st_flags<EPT>(pv, violated);
    px += static_cast<int64_t>(NX) * n;
    po += static_cast<int64_t>(n_obs) * n;
    pr += n;
    pd += n;
    pv += n;
    pg = pg ? pg + static_cast<int64_t>(n_con) * n : pg;

    if (more) {                                                                // hand interval k + R to the integrator
#pragma unroll
      for (int j = 0; j < NU; ++j) sp.a[s][j][lane] = make_double2(a_pre[j][0], a_pre[j][1]);
      mbar_arrive(&sp.full_a[s]);
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e)
    if (io.done_state) io.done_state[env0 + e] = done[e] ? 1 : 0;
}

}  // namespace pcg
