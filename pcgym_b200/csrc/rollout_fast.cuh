// Straight-line specialisation of the control-interval kernel for the common configuration.
//
// Why it exists (profiles/README.md): at 65 536 envs a B200 sub-partition holds only ~110 envs, i.e. four
// dependency chains.  Each env-step is a serial chain of ~650 instructions at ~8 cycles per dependent
// instruction (DFMA latency 8, measured), so the generic kernel runs at IPC ~0.5 and its time is the LENGTH of
// that chain, not the FP64 or HBM throughput.  The only parallelism left is between the phases of one
// env-step: the Gaussian draw (depends on (seed, env, t) only), the integrator, and the bookkeeping of the
// PREVIOUS interval (constraints, reward, observation, stores) are mutually independent.  ptxas interleaves
// independent work only inside one basic block, so this kernel
//   * has no branch in the loop body: every option of the generic kernel that survives here is applied
//     arithmetically (scale 1 / offset 0, zero constraint rows, zero reward weights, predicated stores);
//   * is software-pipelined: iteration k integrates interval k while it finishes interval k-1.
// Eligibility is decided on the host (launch.cuh: fast_ok); everything else takes rollout_kernel.
#pragma once
#include <type_traits>

#include "rollout_kernel.cuh"

namespace pcg {

constexpr int kFastNsp = 2;    // setpoints
constexpr int kFastNcon = 2;   // constraints

template <class M, int EPT>
struct StepOut {               // what interval k leaves for its deferred epilogue
  double uk[EPT][M::NU + M::ND];
  double z[EPT][M::NX];
  double sp_cur[kFastNsp][EPT], sp_nxt[kFastNsp][EPT];
  int t_new;
};

constexpr int kFastPf = 8;          // action-row prefetch distance in control intervals (ring slots per thread)

template <int EPT>
PCG_DEV void cp_async_row(double* smem, const double* gmem) {      // EPT doubles of this thread, global -> shared
  const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  if constexpr (EPT == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
PCG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
PCG_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class M, class TB, int NOISE, int EPT>
__global__ void __launch_bounds__(kBlock, EPT == 1 ? M::MIN_BLOCKS_EPT1 : M::MIN_BLOCKS)
fast_rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io,
                    const __grid_constant__ Derived<M, TB> dv, const int64_t n, const int t0, const int k_steps) {
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND;
  __shared__ __align__(16) double icdf_s[NOISE == PCG_NOISE_PHILOX ? kIcdfDoubles : 2];
  if constexpr (NOISE == PCG_NOISE_PHILOX) {
    fill_icdf_tab(icdf_s);
    __syncthreads();
  }
  // Action rows are staged through shared memory kFastPf intervals ahead with cp.async (global -> shared without a
  // register in between; every thread reads back only the bytes it requested, so no barrier is involved): at
  // 65 536 envs an SM holds ~7 warps of this kernel and a prefetch distance of ONE interval left every interval
  // waiting for most of a DRAM round trip (ncu r2_fast_first_order_65k: long-scoreboard 2.9 of 4.0 stall cycles per
  // issue, 34 % of the HBM roofline for the trivial first-order RHS).
  __shared__ __align__(16) double a_ring[kFastPf][NU][kBlock * EPT];
  const int64_t env0 = (static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x) * EPT;
  if (env0 >= n) return;

  const int flags = d.flags;
  const bool done_on_v = flags & PCG_F_DONE_ON_VIOLATION;
  const int nsp = d.nsp, n_con = d.n_con;
  const int nsp_obs = (flags & PCG_F_SP_HIDDEN) ? 0 : nsp, n_obs = NX + nsp_obs;
  const double pen = d.penalty;
  const PhiloxKeys& key = dv.rk;
  const int t_last = d.N - 1;

  double x[EPT][NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v[EPT];
    ld_row<EPT>(io.x + i * n + env0, v);
#pragma unroll
    for (int e = 0; e < EPT; ++e) x[e][i] = v[e];
  }
  bool done[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) done[e] = (io.done_state && !(flags & PCG_F_EPISODE_START)) ? (io.done_state[env0 + e] != 0) : false;

  const double* pa = io.a + env0;
  double* px = io.x_traj + env0;
  double* po = io.obs + env0;
  double* pr = io.reward + env0;
  double* pg = io.g ? io.g + env0 : nullptr;
  uint8_t* pd = io.done + env0;
  uint8_t* pv = io.violated + env0;

  // setpoint schedule pointers (rows >= nsp alias row 0: loaded, never used)
  const double* psp[kFastNsp];
  double sp_cur[kFastNsp][EPT], sp_nxt[kFastNsp][EPT];
#pragma unroll
  for (int j = 0; j < kFastNsp; ++j) {
    const int jj = j < nsp ? j : 0;
    psp[j] = io.sp + jj * io.sp_sj + env0 * io.sp_se + static_cast<int64_t>(t0) * io.sp_st;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      sp_cur[j][e] = psp[j][e * io.sp_se];
      sp_nxt[j][e] = psp[j][e * io.sp_se + io.sp_st];
    }
    psp[j] += io.sp_st;                                       // -> SP[t0 + 1]; advanced before each use
  }

  auto to_uk = [&](const double (&a_raw)[NU][EPT], double (&uk)[EPT][NUK]) {   // pcgym.py:372-375, 386-414
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int j = 0; j < NU; ++j) uk[e][j] = fma(a_raw[j][e], d.a_scale[j], d.a_off[j]);
#pragma unroll
      for (int i = 0; i < ND; ++i) uk[e][NU + i] = d.d_default[i];
    }
  };
  // g = G [x; uk] + h, dense; rows >= n_con are zero-filled by the host so they can never flag
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
#pragma unroll
        for (int i = 0; i < NUK; ++i) acc = fma(d.con_gu[c][i], uk[e][i], acc);
        g[e] = acc;
        viol[e] = viol[e] || (acc > 0.0);
      }
      if (gout != nullptr && c < n_con) st_row<EPT>(gout + c * n, g);
    }
  };

  auto stage = [&](int k) {                        // action rows of interval k -> ring slot k % kFastPf (one group)
    if (k < k_steps) {
#pragma unroll
      for (int j = 0; j < NU; ++j)
        cp_async_row<EPT>(&a_ring[k % kFastPf][j][threadIdx.x * EPT], pa + (static_cast<int64_t>(k) * NU + j) * n);
    }
    cp_async_commit();                             // committed even when empty: interval k is always group k
  };
#pragma unroll
  for (int k = 0; k < kFastPf; ++k) stage(k);
  double a_nxt[NU][EPT];                           // interval 0 (also feeds the t == 0 pre-check)
  cp_async_wait<kFastPf - 1>();
#pragma unroll
  for (int j = 0; j < NU; ++j)
#pragma unroll
    for (int e = 0; e < EPT; ++e) a_nxt[j][e] = a_ring[0][j][threadIdx.x * EPT + e];

  // ---- t == 0 pre-check (pcgym.py:416-420) with the first interval's input, outside the pipelined loop
  if (t0 == 0 && n_con > 0) {
    double uk0[EPT][NUK];
    to_uk(a_nxt, uk0);
    bool v0[EPT];
    constraints(x, uk0, io.g0 ? io.g0 + env0 : nullptr, v0);
#pragma unroll
    for (int e = 0; e < EPT; ++e) done[e] = done[e] || (v0[e] && done_on_v);
  }

  // finish interval `o` given its end state xe: constraints, done, reward, noisy observation, stores
  auto epilogue = [&](const StepOut<M, EPT>& o, const double (&xe)[EPT][NX]) {
    bool violated[EPT];
    constraints(xe, o.uk, pg, violated);                                       // pcgym.py:444-446
    const bool last = o.t_new == t_last;                                       // pcgym.py:448-449
    double r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      done[e] = done[e] || (violated[e] && done_on_v) || last;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < kFastNsp; ++j) {                                     // r_scale[j >= nsp] == 0
        const double err = pick<NX>(xe[e], d.sp_state[j]) - o.sp_nxt[j][e];
        acc = fma(-(err * err), d.r_scale[j], acc);                            // pcgym.py:555
        acc -= (violated[e] && j < nsp) ? pen : 0.0;                           // once per SP key (pcgym.py:556)
      }
      r[e] = acc;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v[EPT], w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        v[e] = xe[e][i];
        double ox = (NOISE != PCG_NOISE_OFF) ? fma(o.z[e][i] * xe[e][i], d.noise_pct[i], xe[e][i]) : xe[e][i];
        w[e] = fma(ox, d.o_scale[i], d.o_off[i]);          // pcgym.py:483-493, 495-498 (folded: fold_desc)
      }
      st_row<EPT>(px + i * n, v);
      st_row<EPT>(po + i * n, w);
    }
#pragma unroll
    for (int j = 0; j < kFastNsp; ++j) {
      double w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e)
        w[e] = fma(o.sp_cur[j][e], d.o_scale[NX + j], d.o_off[NX + j]);
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
  };

  StepOut<M, EPT> prev;
  double x_prev_end[EPT][NX];
  // one pipelined interval; WITH_EPI is a compile-time flag so the first (peeled) iteration has no branch
  auto interval = [&](const int k, auto with_epi) {
    const int t = t0 + k;
    StepOut<M, EPT> cur;
    cur.t_new = t + 1;                                                         // pcgym.py:441
    double a_raw[NU][EPT];
    cp_async_wait<kFastPf - 1>();                                              // group k (this interval's rows) has landed
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      if constexpr (EPT == 2) {
        const double2 v = *reinterpret_cast<const double2*>(&a_ring[k % kFastPf][j][threadIdx.x * 2]);
        a_raw[j][0] = v.x;
        a_raw[j][1] = v.y;
      } else {
        a_raw[j][0] = a_ring[k % kFastPf][j][threadIdx.x];
      }
    }
    to_uk(a_raw, cur.uk);
    stage(k + kFastPf);                                                        // refill the slot just read (same thread)
    // next interval's setpoints (clamped instead of branching at the end)
#pragma unroll
    for (int j = 0; j < kFastNsp; ++j) {
      psp[j] += (t + 2 < d.N) ? io.sp_st : 0;                                  // -> SP[min(t + 2, N - 1)]
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        cur.sp_cur[j][e] = sp_cur[j][e];
        cur.sp_nxt[j][e] = sp_nxt[j][e];
        sp_cur[j][e] = sp_nxt[j][e];
        sp_nxt[j][e] = psp[j][e * io.sp_se];
      }
    }

    // Gaussian draws of this interval: a pure function of (seed, env, t)
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int i = 0; i < NX; ++i) cur.z[e][i] = 0.0;
    if constexpr (NOISE != PCG_NOISE_OFF) {
      const uint64_t gid0 = static_cast<uint64_t>(d.env_id0 + env0);           // even: launch.cuh checks env_id0
#pragma unroll
      for (int i = 0; i < NX; i += 2) {
        double za[EPT], zb[EPT];
        draw_state_pair<EPT, NOISE == PCG_NOISE_PHILOX_F32>(gid0, static_cast<uint32_t>(cur.t_new), i / 2, key, icdf_s, za, zb);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          cur.z[e][i] = za[e];
          if (i + 1 < NX) cur.z[e][i + 1] = zb[e];
        }
      }
    }

    // interval k-1 is finished from a copy of its end state while interval k is integrated: the two are
    // independent, and with no branch between them they share one basic block
    if constexpr (decltype(with_epi)::value) epilogue(prev, x_prev_end);
    typename M::Aux aux[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) M::prep(dv.p, cur.uk[e], aux[e]);
    rk_step_fixed<M, TB, EPT>(dv, aux, x, cur.uk);
    prev = cur;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int i = 0; i < NX; ++i) x_prev_end[e][i] = x[e][i];
  };
  interval(0, std::false_type{});
  // unrolled by two so the hand-over registers (prev / x_prev_end) alternate names instead of being copied
#pragma unroll 1
  for (int k = 1; k < k_steps; ++k) interval(k, std::true_type{});
  epilogue(prev, x_prev_end);

  unsigned status[EPT];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      v[e] = x[e][i];
      if (i == 0) status[e] = 0;
      if (!isfinite(v[e])) status[e] |= PCG_ST_NONFINITE;
    }
    st_row<EPT>(io.x + i * n + env0, v);
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    if (io.done_state) io.done_state[env0 + e] = done[e] ? 1 : 0;
    if (io.status && status[e]) io.status[env0 + e] |= static_cast<uint8_t>(status[e]);
    if (io.n_rhs) io.n_rhs[env0 + e] += k_steps * TB::S;
  }
}

}  // namespace pcg
