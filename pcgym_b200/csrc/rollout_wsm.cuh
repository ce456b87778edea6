// One persistent CTA per SM, warp-specialised, with DYNAMICALLY balanced finisher work.
//
// Why it exists (profiles/README.md): rollout_ws.cuh ties each finisher warp to one integrator warp.  At the headline
// batch (65 536 envs = 6.92 tiles of 64 envs per SM) seven such pairs land on four warp schedulers as 4+4+3+3 warps,
// the two schedulers that carry two integrator warps AND two finisher warps set the pace, and the kernel loses
// ~25 % to that quantisation.  Here a CTA of 16 warps owns everything its SM has to do (up to 8 tiles):
//   * warps 0..NI-1 integrate one tile each (serial in time; identical to the integrator role of rollout_ws.cuh);
//   * the other 16-NI warps are a POOL of finishers.  The finishing work of tile p at interval k -- Gaussian draws,
//     constraints, reward, observation, all global traffic, and the prefetch of the action rows of interval k+R --
//     is a task q = k*NI + p that any finisher takes from a shared counter.  Finishers on a scheduler that carries
//     only one integrator warp get more issue slots, finish their tasks sooner and therefore take more of them:
//     the load evens out across the four schedulers without any static partition.
// A task is stateless given (p, k) as long as `done` does not depend on earlier violations, so the kernel requires
// done_on_cons_vio == False (the launcher falls back otherwise).  Same ring / mbarrier hand-over as rollout_ws.cuh.
//
// Round 2 (profiles/README.md "r2"): the ncu source page showed the kernel bound by LATENCY, not by issue slots or the
// FP64 pipe (both ~50 % busy): 16 resident warps, each a dependent chain with ILP ~2, cannot cover 8-cycle DFMA
// dependencies, ~30-cycle shared-memory and ~1000-cycle global loads.  Removing a quarter of the FP64 instructions
// changed the time by 4 %.  So the integrators now carry ONE env per lane and there are twice as many of them
// (two integrator warps per 64-env tile, lane L of warp h owns env 2L + h of the tile): 14 + 6 = 20 warps of <= 96
// registers per SM at the headline batch, 3.5 integrator chains per scheduler instead of 1.75; and the action rows
// reach the ring by cp.async (global -> shared, completion signalled on the tile's mbarrier), so no finisher ever
// waits on a global load.
#pragma once
#include "rollout_ws.cuh"

namespace pcg {

#ifndef PCG_WSM_EPI
#define PCG_WSM_EPI 1              // envs per integrator lane (2 = the round-1 layout, kept for A/B)
#endif
// Warp layout of the CTA.  With PCG_WSM_INT_REGS != 0 the first 16 warps (four warpgroups) are the integrator class
// and give registers back (setmaxnreg.dec) which the finisher warpgroups behind them take (setmaxnreg.inc): ptxas
// needs 54 registers for the one-env integrator and ~74-92 for the finisher, so the register file holds
//   24 warps: 16 x 56 + 8 x 96        28 warps: 16 x 56 + 12 x 88        32 warps: 16 x 56 + 16 x 72
// With PCG_WSM_INT_REGS == 0 every warp keeps the launch budget, integrators first, the rest finishers (A/B: 20 warps x 96).
#ifndef PCG_WSM_WARPS
#define PCG_WSM_WARPS 28
#endif
#ifndef PCG_WSM_INT_REGS
#define PCG_WSM_INT_REGS 56
#endif
#ifndef PCG_WSM_FIN_REGS
#define PCG_WSM_FIN_REGS 88
#endif
#ifndef PCG_WSM_RING
#define PCG_WSM_RING 16            // ring slots per tile (8 makes room for CstrW's direct exponential table: A/B)
#endif
#ifndef PCG_WSM_PREFETCH
#define PCG_WSM_PREFETCH 1         // take the next task id while the current task runs (hides the shared-memory atomic)
#endif
constexpr int kWsmEpi = PCG_WSM_EPI;
constexpr int kWsmIpt = 2 / kWsmEpi;                       // integrator warps per 64-env tile
constexpr int kWsmWarps = kWsmEpi == 2 ? 16 : PCG_WSM_WARPS;
constexpr int kWsmIntClass = (kWsmEpi == 1 && PCG_WSM_INT_REGS != 0) ? 16 : 0;   // warps [0, 16) hold the small register budget
constexpr int kWsmLaunchRegs = (65536 / (32 * kWsmWarps)) / 8 * 8;
static_assert(kWsmIntClass == 0 ||
              16 * (kWsmLaunchRegs - PCG_WSM_INT_REGS) >= (kWsmWarps - 16) * (PCG_WSM_FIN_REGS - kWsmLaunchRegs),
              "setmaxnreg budgets do not fit the register file");
constexpr int kWsmMaxTiles = 8;    // 64-env tiles per CTA (<= 512 envs per SM)
static_assert(kWsmIntClass == 0 || kWsmIntClass >= kWsmMaxTiles * kWsmIpt, "integrator class must hold every integrator warp");
// Ring depth.  An mbarrier wait names its phase by PARITY only, so a waiter may never be a whole phase ahead of the
// barrier.  A finisher holding interval k + RING of a tile while interval k of the same tile is still waiting for the
// integrator would be exactly that.  The task loop rules it out with a gate on full_a (see there): a finisher looks at
// full_x[s] for interval k only after task (p, k - RING) has completed.  The gate's own wait is exact while a tile has
// at most 2 RING tasks outstanding, which RING >= the number of finishers (capped at kWsmMaxFin; each owns its current
// and its prefetched task) guarantees.
constexpr int kWsmRing = PCG_WSM_RING;
constexpr int kWsmMaxFin = PCG_WSM_RING;
static_assert(kWsmRing >= kWsmMaxFin && (kWsmRing & (kWsmRing - 1)) == 0, "ring depth vs finisher count");

// ADAPT = the integrator warps run the embedded pair with per-lane step-size control (advance_adaptive_aux: make_env's
// default scheme).  The controller keeps six stage derivatives and two candidate states live and an interval costs
// ~2.5 fixed steps, so the finishers have 2.5x less to do per unit time: 20 warps x 96 registers, 16 integrator-class
// warps + 4 finishers, no register hand-over (setmaxnreg moves registers only inside the CTA's own allocation).
// GEN = the GENERAL variant (BASELINE config 5's cstr part and the paper scripts' disturbance runs): per-interval model
// disturbances travel through the ring as extra rows beside the action rows (dist[t + 1]; per-env tables by 16-byte
// cp.async, the reference's shared schedule by two 8-byte ones), observed disturbance slots are appended to the
// observation, and the integrator warps loop over n_sub fixed substeps.  Two more rows per slot do not fit 16-slot rings
// (8 tiles x 16 slots x 5 rows x 512 B = 320 KB), so GEN runs 8-slot rings with 24 warps (16 x 64 + 8 x 96 registers).
template <bool ADAPT, bool GEN = false>
struct WsmLayout {
  static constexpr int ring = GEN ? 8 : kWsmRing;
  static constexpr int max_fin = ring;                         // see "Ring depth" above
  static constexpr int warps = ADAPT ? 20 : (GEN ? 24 : kWsmWarps);
  static constexpr int launch_regs = (65536 / (32 * warps)) / 8 * 8;
  static constexpr int int_regs = ADAPT ? launch_regs : (GEN ? 64 : (PCG_WSM_INT_REGS ? PCG_WSM_INT_REGS : launch_regs));
  static constexpr int fin_regs = ADAPT ? launch_regs : (GEN ? 96 : (PCG_WSM_INT_REGS ? PCG_WSM_FIN_REGS : launch_regs));
  static_assert(16 * (launch_regs - int_regs) >= (warps - 16) * (fin_regs - launch_regs),
                "setmaxnreg budgets do not fit the CTA's register allocation");
  static_assert(!(ADAPT && GEN), "the general variant is fixed-step");
};

template <int N> PCG_DEV void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> PCG_DEV void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

PCG_DEV void cp_async16(void* smem, const void* gmem) {          // 16 bytes global -> shared, no register staging
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
PCG_DEV void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
// arrive on `b` once every cp.async this thread issued before has landed (counts as one of the expected arrivals)
PCG_DEV void cp_async_arrive(uint64_t* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

constexpr int kWsmSpMax = 256;     // staged setpoint schedule: k_steps + 1 <= kWsmSpMax time indices per setpoint

// NOISE / SPS decide which tables the CTA stages: the inverse-normal-CDF coefficients (16 KB) only when the finishers
// draw in fp64, the setpoint schedule only when it is shared by all envs (the reference's form: io.sp_se == 0)
// ring of one 64-env tile: as WsPair, with NDR extra input rows per slot (GEN: the model's disturbance channels)
template <class M, int RING, int NDR>
struct WsmTile {
  double2 a[RING][M::NU + NDR][32];
  double2 x[RING][M::NX][32];
  uint64_t full_a[RING];
  uint64_t full_x[RING];
};

template <class M, int NOISE, int NSP, bool SPS, int RING = kWsmRing, int NDR = 0>
struct WsmShared {
  WsmTile<M, RING, NDR> tile[kWsmMaxTiles];
  uchar2 done0[kWsmMaxTiles][32];   // sticky done flag each env entered the launch with
  int next_task;
  double exp_tab[exp_tab_doubles<M>()];               // 2^(j/256) for exp_tab_split, or the direct table of exp_wide_split (integrator warps)
  alignas(16) double icdf[NOISE == PCG_NOISE_PHILOX ? kIcdfDoubles : 2];   // kIcdfTab (finishers; read as double2)
  double sp[SPS ? NSP : 1][SPS ? kWsmSpMax : 1];      // SP[j][t0 + i], i = 0..k_steps
};

// NSP = number of setpoints (1 or 2), CONU = some constraint reads an input, SPS = the setpoint schedule is shared by
// all envs and short enough to be staged in shared memory: compile-time, so the finisher carries no dead reward /
// setpoint / input-column work and no branch for it
template <class M, class TB, int NOISE, int NSP, bool CONU, bool SPS, bool ADAPT = false, bool GEN = false>
__global__ void __launch_bounds__(32 * WsmLayout<ADAPT, GEN>::warps, 1)
wsm_rollout_kernel(const __grid_constant__ pcg_desc d, const __grid_constant__ pcg_io io,
                   const __grid_constant__ Derived<M, TB> dv, const int64_t n, const int t0, const int k_steps,
                   const int64_t envs_per_cta, const int n_tiles) {
  using L = WsmLayout<ADAPT, GEN>;
  constexpr int NX = M::NX, NU = M::NU, ND = M::ND, NUK = NU + ND, EPT = 2, R = L::ring, NDR = GEN ? ND : 0;
  using Tile = WsmTile<M, R, NDR>;
  extern __shared__ __align__(16) unsigned char wsm_raw[];
  using Shared = WsmShared<M, NOISE, NSP, SPS, R, NDR>;
  Shared& sh = *reinterpret_cast<Shared*>(wsm_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t cta_base = static_cast<int64_t>(blockIdx.x) * envs_per_cta;
  const int64_t cta_end = (cta_base + envs_per_cta < n) ? cta_base + envs_per_cta : n;

  // lanes that own envs in tile p (the last tile of a CTA may be partial; envs_per_cta and n are even)
  const int cta_pairs = static_cast<int>((cta_end - cta_base) / 2);            // env pairs (= lanes) this CTA owns
  auto tile_lanes = [&](int p) {
    const int left = cta_pairs - p * 32;
    return left < 0 ? 0 : (left > 32 ? 32 : left);
  };
  constexpr int EPI = kWsmEpi, IPT = kWsmIpt;
  const int n_int = n_tiles * IPT;                                             // integrator warps of this CTA
  if (warp < n_tiles && lane == 0) {
    const int cnt = tile_lanes(warp) > 0 ? tile_lanes(warp) : 1;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mbar_init(&sh.tile[warp].full_a[s], cnt);                                // one cp.async arrival per finisher lane
      mbar_init(&sh.tile[warp].full_x[s], cnt * IPT);                          // one arrival per integrator lane
    }
  }
  if (threadIdx.x == 0) sh.next_task = 0;
  fill_exp_tab<M>(sh.exp_tab, dv.p);
  if constexpr (NOISE == PCG_NOISE_PHILOX) {
    fill_icdf_tab(sh.icdf);
  }
  if constexpr (SPS) {               // SP[j][t0 .. t0 + k_steps] (t0 + k_steps <= N - 1: abi.cu validates)
    for (int i = threadIdx.x; i < NSP * (k_steps + 1); i += blockDim.x) {
      const int j = i / (k_steps + 1), c = i - j * (k_steps + 1);
      sh.sp[j][c] = io.sp[j * io.sp_sj + static_cast<int64_t>(t0 + c) * io.sp_st];
    }
  }
  __syncthreads();

  // a_raw rows [0, NU) are the action rows, rows [NU, NU + NDR) the supplied disturbance values of the interval (GEN)
  auto to_uk = [&](const double (&a_raw)[NU + NDR][EPT], double (&uk)[EPT][NUK]) {   // pcgym.py:372-375, 386-414
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
#pragma unroll
      for (int j = 0; j < NU; ++j) uk[e][j] = fma(a_raw[j][e], d.a_scale[j], d.a_off[j]);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        double v = d.d_default[i];
        if constexpr (GEN) v = d.d_src[i] >= 0 ? a_raw[NU + i][e] : v;
        uk[e][NU + i] = v;
      }
    }
  };
  // GEN: stage the supplied disturbance rows of time index tt (= t + 1 of the interval, pcgym.py:386-412) for this lane's
  // env pair into ring slot s of tile `sp`.  Per-env tables are contiguous along the env axis (16-byte copies; the
  // launcher checks alignment), the reference's shared schedule is one value for every env (two 8-byte copies).
  auto stage_dist = [&](Tile& sp, int s, int lane_, int64_t env0_, int tt) {
    if constexpr (GEN) {
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        if (d.d_src[i] >= 0) {
          const double* src = io.dist + static_cast<int64_t>(tt) * io.d_st + d.d_src[i] * io.d_sj + env0_ * io.d_se;
          if (io.d_se != 0) {
            cp_async16(&sp.a[s][NU + i][lane_], src);
          } else {
            cp_async8(&sp.a[s][NU + i][lane_].x, src);
            cp_async8(&sp.a[s][NU + i][lane_].y, src);
          }
        }
      }
    }
  };

  // first finisher warp / number of finishers (warps between the two roles, and past the last finisher, retire)
  const int fin0 = kWsmIntClass ? kWsmIntClass : n_int;
  static_assert(!ADAPT || (EPI == 1 && TB::EMBEDDED), "adaptive integrator warps carry one env per lane");
  const int n_fin = (L::warps - fin0) < L::max_fin ? (L::warps - fin0) : L::max_fin;
  if (warp < fin0) {
    if constexpr (kWsmIntClass != 0 && L::int_regs < L::launch_regs) setmaxnreg_dec<L::int_regs>();   // whole warpgroups: warps 0..15
    if (warp >= n_int) return;
    // ================================================================ integrator warp: tile p, env 2 lane + h (+ e)
    const int p = warp / IPT, h = (warp % IPT) * EPI;
    if (lane >= tile_lanes(p)) return;
    Tile& sp = sh.tile[p];
    const int64_t env0 = cta_base + static_cast<int64_t>(p) * 64 + 2 * lane + h;
    double x[EPI][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v[EPI];
      ld_row<EPI>(io.x + i * n + env0, v);
#pragma unroll
      for (int e = 0; e < EPI; ++e) x[e][i] = v[e];
    }
    // adaptive mode: the step size carried from interval to interval, RHS evaluations and status bits of this lane's env
    double h_carry = (ADAPT && io.h_carry) ? io.h_carry[env0] : 0.0;
    int n_rhs = 0;
    unsigned st_adapt = 0;
    unsigned st_range[EPI] = {};
    // the finishers read the initial state for the t == 0 pre-check before they release interval 0, so io.x may be
    // overwritten at the end without a race
#pragma unroll 1
    for (int k = 0; k < k_steps; ++k) {
      const int s = k & (R - 1);
      mbar_wait_backoff(&sp.full_a[s], (k / R) & 1);
      double uk[EPI][NUK];
#pragma unroll
      for (int e = 0; e < EPI; ++e) {                                          // pcgym.py:372-375, 386-414
#pragma unroll
        for (int j = 0; j < NU; ++j)
          uk[e][j] = fma(reinterpret_cast<const double*>(&sp.a[s][j][lane])[h + e], d.a_scale[j], d.a_off[j]);
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          double v = d.d_default[i];
          if constexpr (GEN) v = d.d_src[i] >= 0 ? reinterpret_cast<const double*>(&sp.a[s][NU + (GEN ? i : 0)][lane])[h + e] : v;
          uk[e][NU + i] = v;
        }
      }
      typename M::Aux aux[EPI];
#pragma unroll
      for (int e = 0; e < EPI; ++e) {
        M::prep(dv.p, uk[e], aux[e]);
        bind_exp_tab<M>(aux[e], sh.exp_tab);
      }
      if constexpr (ADAPT) {                                                   // integrator.py:65-88 (Tsit5 + PID 1e-8)
        advance_adaptive_aux<M, TB>(dv, dv.p, aux[0], x[0], uk[0], d.dt, d.rtol, d.atol, d.max_steps, h_carry, n_rhs, st_adapt);
      } else if constexpr (GEN) {
#pragma unroll 1
        for (int sub = 0; sub < d.n_sub; ++sub) rk_step_fixed<M, TB, EPI>(dv, aux, x, uk);
      } else {
        rk_step_fixed<M, TB, EPI>(dv, aux, x, uk);                             // integrator.py:90-107
      }
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int e = 0; e < EPI; ++e) reinterpret_cast<double*>(&sp.x[s][i][lane])[h + e] = x[e][i];
      mbar_arrive(&sp.full_x[s]);
#pragma unroll
      for (int e = 0; e < EPI; ++e) st_range[e] |= range_flag<M>(dv.p, x[e]);  // wide-table models: state left the table's band
    }
#pragma unroll
    for (int e = 0; e < EPI; ++e) {
      unsigned status = st_adapt | st_range[e];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        if (!isfinite(x[e][i])) status |= PCG_ST_NONFINITE;
        io.x[i * n + env0 + e] = x[e][i];
      }
      if (io.status && status) io.status[env0 + e] |= static_cast<uint8_t>(status);
      if (io.n_rhs) io.n_rhs[env0 + e] += ADAPT ? n_rhs : k_steps * TB::S * (GEN ? d.n_sub : 1);
    }
    if (ADAPT && io.h_carry) io.h_carry[env0] = h_carry;
    return;
  }

  // ================================================================== finisher pool
  if constexpr (kWsmIntClass != 0 && L::fin_regs > L::launch_regs) setmaxnreg_inc<L::fin_regs>();                       // whole warpgroups: warps 16 and up
  const int fin = warp - fin0;
  if (fin >= n_fin) return;
  const int flags = d.flags;
  const int nsp = d.nsp, n_con = d.n_con;
  const int nsp_obs = (flags & PCG_F_SP_HIDDEN) ? 0 : nsp, nd_obs = GEN ? d.nd_obs : 0, n_obs = NX + nsp_obs + nd_obs;
  const double pen = d.penalty;
  const PhiloxKeys& key = dv.rk;
  const int t_last = d.N - 1;

  // g = G [x; uk] + h, dense; rows >= n_con are zero-filled by the host so they can never flag
  // (addresses in the finisher are 32-bit element indices off the array bases -- one IMAD.WIDE each; wsm_ok()
  // checks that every array of the launch has fewer than 2^31 elements)
  const uint32_t n32 = static_cast<uint32_t>(n);
  auto constraints = [&](const double (&xs)[EPT][NX], const double (&uk)[EPT][NUK], double* gout, uint32_t gidx,
                         bool (&viol)[EPT]) {
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
        if constexpr (CONU) {                            // most constraints read states only
#pragma unroll
          for (int i = 0; i < NUK; ++i) acc = fma(d.con_gu[c][i], uk[e][i], acc);
        }
        g[e] = acc;
        viol[e] = viol[e] || (acc > 0.0);
      }
      if (gout != nullptr && c < n_con) st_row<EPT>(gout + (gidx + c * n32), g);
    }
  };

  // ---- prologue, tiles dealt round-robin to the finishers: t == 0 pre-check on the initial state, then the
  // first R intervals' action rows into the ring (which releases the tile's integrator)
#pragma unroll 1
  for (int p = fin; p < n_tiles; p += n_fin) {
    if (lane >= tile_lanes(p)) continue;
    Tile& sp = sh.tile[p];
    const int64_t env0 = cta_base + static_cast<int64_t>(p) * 64 + 2 * lane;
    const double* pa = io.a + env0;
    double a0[NU + NDR][EPT];
#pragma unroll
    for (int j = 0; j < NU; ++j) ld_row<EPT>(pa + j * n, a0[j]);
    if constexpr (GEN) {                                                       // dist[t0 + 1] for the pre-check's input vector
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        a0[NU + i][0] = a0[NU + i][1] = 0.0;
        if (d.d_src[i] >= 0) {
          const double* src = io.dist + static_cast<int64_t>(t0 + 1) * io.d_st + d.d_src[i] * io.d_sj + env0 * io.d_se;
          a0[NU + i][0] = src[0];
          a0[NU + i][1] = src[io.d_se];
        }
      }
    }
    sh.done0[p][lane] = (io.done_state && !(d.flags & PCG_F_EPISODE_START)) ? *reinterpret_cast<const uchar2*>(io.done_state + env0)
                                                                             : make_uchar2(0, 0);
    if (t0 == 0 && n_con > 0) {                                                // pcgym.py:416-420
      double x0[EPT][NX], uk0[EPT][NUK];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v[EPT];
        ld_row<EPT>(io.x + i * n + env0, v);
        x0[0][i] = v[0];
        x0[1][i] = v[1];
      }
      to_uk(a0, uk0);
      bool v0[EPT];
      constraints(x0, uk0, io.g0, static_cast<uint32_t>(env0), v0);
    }
#pragma unroll
    for (int s = 0; s < R; ++s) {
      if (s < k_steps) {
#pragma unroll
        for (int j = 0; j < NU; ++j) cp_async16(&sp.a[s][j][lane], pa + (static_cast<int64_t>(s) * NU + j) * n);
        stage_dist(sp, s, lane, env0, t0 + s + 1);
        cp_async_arrive(&sp.full_a[s]);
      }
    }
  }

  // ---- task loop
  const int n_tasks = k_steps * n_tiles;
  // floor(q / n_tiles) = umulhi(q, ceil(2^32 / n_tiles)) for q < 2^32 / n_tiles (n_tiles <= 8, q < 2^27: launch.cuh)
  const uint32_t tiles_rcp = static_cast<uint32_t>((uint64_t{1} << 32) / static_cast<uint32_t>(n_tiles)) + 1u;
#pragma unroll 1
#if PCG_WSM_PREFETCH
  int q_nxt = 0;
  if (lane == 0) q_nxt = atomicAdd(&sh.next_task, 1);
#endif
  while (true) {
#if PCG_WSM_PREFETCH
    const int q = __shfl_sync(0xffffffffu, q_nxt, 0);
    if (q >= n_tasks) break;
    if (lane == 0) q_nxt = atomicAdd(&sh.next_task, 1);                        // consumed one iteration later
#else
    int q = 0;
    if (lane == 0) q = atomicAdd(&sh.next_task, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= n_tasks) break;
#endif
    const int k = static_cast<int>(__umulhi(static_cast<uint32_t>(q), tiles_rcp)), p = q - k * n_tiles;   // q / n_tiles
    if (lane >= tile_lanes(p)) continue;
    Tile& sp = sh.tile[p];
    const int s = k & (R - 1);
    const int t = t0 + k, t_new = t + 1;                                       // pcgym.py:441
    const uint32_t env0 = static_cast<uint32_t>(cta_base) + static_cast<uint32_t>(p) * 64u + 2u * lane;
    const uint32_t kn = static_cast<uint32_t>(k) * n32;

    const bool more = k + R < k_steps;

    // SP[t] (shown in the observation) and SP[t+1] (enters the reward)
    double spc[NSP][EPT], spn[NSP][EPT];
    if constexpr (SPS) {                                    // one value for the whole warp, from shared memory
#pragma unroll
      for (int j = 0; j < NSP; ++j) {
        const double c = sh.sp[j][k], nx = sh.sp[j][k + 1];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          spc[j][e] = c;
          spn[j][e] = nx;
        }
      }
    } else {
      const uint32_t sp_st = static_cast<uint32_t>(io.sp_st), sp_se = static_cast<uint32_t>(io.sp_se);
#pragma unroll
      for (int j = 0; j < NSP; ++j) {
        const uint32_t si = j * static_cast<uint32_t>(io.sp_sj) + env0 * sp_se + static_cast<uint32_t>(t) * sp_st;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          spc[j][e] = io.sp[si + e * sp_se];
          spn[j][e] = io.sp[si + e * sp_se + sp_st];
        }
      }
    }

    // Gaussian draws of this interval: a pure function of (seed, env, t), computed while the integrator works
    double z[EPT][NX];
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int i = 0; i < NX; ++i) z[e][i] = 0.0;
    if constexpr (NOISE != PCG_NOISE_OFF) {
      const uint64_t gid0 = static_cast<uint64_t>(d.env_id0) + env0;            // even: launch.cuh checks env_id0
#pragma unroll
      for (int i = 0; i < NX; i += 2) {
        double za[EPT], zb[EPT];
        draw_state_pair<EPT, NOISE == PCG_NOISE_PHILOX_F32>(gid0, static_cast<uint32_t>(t_new), i / 2, key, sh.icdf, za, zb);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          z[e][i] = za[e];
          if (i + 1 < NX) z[e][i + 1] = zb[e];
        }
      }
    }
    const int c = k / R;
    // Gate: slot s may only be waited on once task (p, k - R) has completed, i.e. once ITS prefetch of this interval's
    // action rows has landed (phase c of full_a[s]).  Without it a finisher could reach interval k while interval k - R
    // of the same tile is still outstanding -- a finisher owns its current AND its prefetched task, so R + 1 outstanding
    // tasks of one tile need only (R + 1) / 2 finishers when that tile's integrators lag (two integrator warps on one
    // scheduler against one for another tile) -- and full_x[s], still a whole phase behind, would let the parity-only
    // wait below pass at once (measured: 20 000 envs = 3 tiles per SM hung with the 8-slot ring).  The gate itself is
    // exact as long as fewer than 2 R + 1 tasks of a tile are owned at a time: n_fin <= R finishers own at most 2 R.
    if (k >= R) mbar_wait_backoff(&sp.full_a[s], c & 1);
    mbar_wait_backoff(&sp.full_x[s], c & 1);
    // (written in the prologue before the tile's ring was released: ordered by the full_a -> full_x chain)
    const uchar2 d0 = sh.done0[p][lane];
    const bool done0[EPT] = {d0.x != 0, d0.y != 0};
    double xe[EPT][NX], uk[EPT][NUK];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const double2 v = sp.x[s][i][lane];
      xe[0][i] = v.x;
      xe[1][i] = v.y;
    }
    if constexpr (CONU || GEN) {                            // the input vector: constraints that read it, observed disturbances
      double a_raw[NU + NDR][EPT];
#pragma unroll
      for (int j = 0; j < NU + NDR; ++j) {
        const double2 v = sp.a[s][j][lane];
        a_raw[j][0] = v.x;
        a_raw[j][1] = v.y;
      }
      to_uk(a_raw, uk);
    }

    const uint32_t row = kn + env0;                                            // element [k][env0] of the [K][n] arrays
    bool violated[EPT], done[EPT];
    constraints(xe, uk, io.g, kn * n_con + env0, violated);                    // pcgym.py:444-446
    const bool last = t_new == t_last;                                         // pcgym.py:448-449
    double r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      done[e] = done0[e] || last;                                              // done_on_cons_vio is off (launcher)
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NSP; ++j) {
        const double err = pick<NX>(xe[e], d.sp_state[j]) - spn[j][e];
        acc = fma(-(err * err), d.r_scale[j], acc);                            // pcgym.py:555
        acc -= violated[e] ? pen : 0.0;                                        // once per SP key (pcgym.py:556)
      }
      r[e] = acc;
    }
    const uint32_t ix = kn * NX + env0, io_ = kn * n_obs + env0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v[EPT], w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        v[e] = xe[e][i];
        const double ox = (NOISE != PCG_NOISE_OFF) ? fma(z[e][i] * xe[e][i], d.noise_pct[i], xe[e][i]) : xe[e][i];
        w[e] = fma(ox, d.o_scale[i], d.o_off[i]);          // pcgym.py:483-493, 495-498 (folded: fold_desc)
      }
      st_row<EPT>(io.x_traj + (ix + i * n32), v);
      st_row<EPT>(io.obs + (io_ + i * n32), w);
    }
#pragma unroll
    for (int j = 0; j < NSP; ++j) {
      double w[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) w[e] = fma(spc[j][e], d.o_scale[NX + j], d.o_off[NX + j]);
      if (j < nsp_obs) st_row<EPT>(io.obs + (io_ + (NX + j) * n32), w);
    }
    if constexpr (GEN) {                                    // observed disturbance slots (pcgym.py:386-412): value held over the interval
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double w[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          double v = 0.0;
#pragma unroll
          for (int i = 0; i < ND; ++i) v = (d.d_obs[j] == i) ? uk[e][NU + i] : v;
          w[e] = fma(v, d.o_scale[NX + nsp_obs + j], d.o_off[NX + nsp_obs + j]);
        }
        if (j < nd_obs) st_row<EPT>(io.obs + (io_ + (NX + nsp_obs + j) * n32), w);
      }
    }
    st_row<EPT>(io.reward + row, r);
    st_flags<EPT>(io.done + row, done);
    st_flags<EPT>(io.violated + row, violated);

    if (more) {                               // interval k + R: global -> ring slot s, arrival when the bytes have landed
#pragma unroll
      for (int j = 0; j < NU; ++j)
        cp_async16(&sp.a[s][j][lane], io.a + ((static_cast<uint32_t>(k + R) * NU + j) * n32 + env0));
      stage_dist(sp, s, lane, static_cast<int64_t>(env0), t + R + 1);
      cp_async_arrive(&sp.full_a[s]);
    }
    if (k == k_steps - 1 && io.done_state) {                                   // the sticky flag after this launch
#pragma unroll
      for (int e = 0; e < EPT; ++e) io.done_state[env0 + e] = done[e] ? 1 : 0;
    }
  }
}

}  // namespace pcg
