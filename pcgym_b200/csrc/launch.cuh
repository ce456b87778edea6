// Host-side launch of rollout_kernel for one model: picks the (scheme, adaptive, general) instantiation,
// computes the derived constant block and the vector-access eligibility.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include "rollout_lanes.cuh"
#include "rollout_wsm.cuh"

namespace pcg {

using launch_fn = cudaError_t (*)(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);

// which kernel family the calling thread launched last (pcg_last_kernel, diagnostics for bench.py)
inline thread_local const char* t_last_kernel = "";

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 128-bit row access needs an even env count (every row then starts 16-byte aligned) and aligned bases
inline bool vector_ok(const pcg_io& io, int64_t n) {
  if (n % 2) return false;
  const void* f64[] = {io.x, io.a, io.a_save, io.z, io.tail, io.x_traj, io.obs, io.reward, io.g, io.g0, io.u_prev, io.theta};
  for (const void* p : f64)
    if (!al16(p)) return false;
  return (reinterpret_cast<uintptr_t>(io.done) & 1u) == 0 && (reinterpret_cast<uintptr_t>(io.violated) & 1u) == 0 &&
         (reinterpret_cast<uintptr_t>(io.done_state) & 1u) == 0;
}

template <class M>
inline bool needs_general(const pcg_desc& d) {
  if (d.n_tail > 0 || d.nd > 0 || d.nsp > 2 || d.n_con > 4) return true;
  for (int c = 0; c < d.n_con; ++c)
    if (d.con_mask_x[c] >> M::NX) return true;
  return false;
}

template <class M, class TB, bool ADAPTIVE, bool GENERAL, int EPT>
cudaError_t launch_one(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const int64_t threads = (n_env + EPT - 1) / EPT;
  const unsigned grid = static_cast<unsigned>((threads + kBlock - 1) / kBlock);
  const Derived<M, TB> dv = make_derived<M, TB>(d);
  rollout_kernel<M, TB, ADAPTIVE, GENERAL, EPT><<<grid, kBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
  t_last_kernel = EPT == 2 ? "rollout_kernel (generic, 2 envs/thread)" : "rollout_kernel (generic, 1 env/thread)";
  return cudaGetLastError();
}

// environments per thread: the model's preference for the lock-step fixed integrator (2 = ILP 2 and
// 128-bit rows), 1 for the adaptive integrator (lanes take different step counts anyway).
// PCG_EPT=1 in the environment forces one env per thread (profiling knob, read once).
inline int ept_override() {
  static const int v = [] { const char* e = std::getenv("PCG_EPT"); return e ? std::atoi(e) : 0; }();
  return v;
}

// the straight-line, software-pipelined kernel covers: fixed step with n_sub == 1, lean observation vector,
// SP reward with <= 2 setpoints, <= 2 constraints, no a_delta, Philox / no noise, all trajectory outputs wanted
template <class M>
inline bool fast_ok(const pcg_desc& d, const pcg_io& io, bool adaptive = false) {
  if (adaptive != ((d.flags & PCG_F_ADAPTIVE) != 0) || (!adaptive && d.n_sub != 1) || needs_general<M>(d)) return false;
  if ((d.flags & PCG_F_A_DELTA) && (d.flags & PCG_F_NORMALISE_A)) return false;
  if (d.reward != PCG_REWARD_SP || d.nsp < 1 || d.nsp > kFastNsp || d.n_con > kFastNcon) return false;
  if (d.noise == PCG_NOISE_EXTERNAL) return false;
  if (!io.x_traj || !io.obs || !io.reward || !io.done || !io.violated) return false;
  if (d.n_con > 0 && !io.g) return false;
  return true;
}

// The straight-line kernels apply every option arithmetically: the descriptor they see has the options folded
// into its coefficients (scale 1 / offset 0 when a normalisation is off, scale 0 / offset 0 for a masked
// observation -- a non-finite masked state then reads NaN instead of 0 --, penalty 0 without r_penalty).
template <class M>
inline pcg_desc fold_desc(const pcg_desc& d) {
  pcg_desc f = d;
  if (!(d.flags & PCG_F_NORMALISE_A))
    for (int j = 0; j < PCG_MAX_NU; ++j) { f.a_scale[j] = 1.0; f.a_off[j] = 0.0; }
  for (int i = 0; i < PCG_MAX_OBS; ++i) {
    if (!(d.flags & PCG_F_NORMALISE_O)) { f.o_scale[i] = 1.0; f.o_off[i] = 0.0; }
    if (i < 32 && (d.obs_zero_mask >> i & 1u)) { f.o_scale[i] = 0.0; f.o_off[i] = 0.0; }
  }
  if (!(d.flags & PCG_F_R_PENALTY)) f.penalty = 0.0;
  return f;
}

template <class M, class TB, int EPT>
cudaError_t launch_fast(const pcg_desc& d0, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const pcg_desc d = fold_desc<M>(d0);
  const int64_t threads = (n_env + EPT - 1) / EPT;
  const unsigned grid = static_cast<unsigned>((threads + kBlock - 1) / kBlock);
  const Derived<M, TB> dv = make_derived<M, TB>(d);
  switch (d.noise) {
    case PCG_NOISE_OFF:
      fast_rollout_kernel<M, TB, PCG_NOISE_OFF, EPT><<<grid, kBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    case PCG_NOISE_PHILOX:
      fast_rollout_kernel<M, TB, PCG_NOISE_PHILOX, EPT><<<grid, kBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    default:
      fast_rollout_kernel<M, TB, PCG_NOISE_PHILOX_F32, EPT><<<grid, kBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
  }
  t_last_kernel = EPT == 2 ? "fast_rollout_kernel (software-pipelined, 2 envs/thread)"
                           : "fast_rollout_kernel (software-pipelined, 1 env/thread)";
  return cudaGetLastError();
}

// warp-specialised kernel (rollout_ws.cuh): models that opt in (M::WS > 0 = pairs per CTA), 128-bit rows.
// PCG_WS=0 forces the monolithic fast kernel, PCG_WS=1 the pair-per-CTA kernel, PCG_WS=3 the per-SM kernel
// (profiling knobs).
inline int ws_override() {
  static const int v = [] { const char* e = std::getenv("PCG_WS"); return e ? std::atoi(e) : -1; }();
  return v;
}

// Which straight-line kernel for a batch (profiles/r1_sweep_envs.txt, 148 SMs): the warp-specialised kernel
// keeps 3-4 light warps per scheduler and wins when the batch fills the SMs evenly or thinly (<= 6 pairs of 64
// envs per SM) and again from ~10 pairs per SM up (16 resident warps against the monolithic kernel's 8); in
// between (the 7th and 8th pair land two-deep on two of the four schedulers) the 255-register kernel is ahead.
// (Where the per-SM kernel of rollout_wsm.cuh is eligible it is taken before this choice is consulted.)
inline int sm_count();
inline bool ws_preferred(int64_t n_env) {
  const int sms = sm_count();
  const double pairs_per_sm = static_cast<double>(n_env) / (64.0 * sms);
  return pairs_per_sm <= 6.0 || pairs_per_sm > 8.5;
}

template <class M, class TB, int W, int NSP>
cudaError_t launch_ws_s(const pcg_desc& d, const pcg_io& io, const Derived<M, TB>& dv, int64_t n_env, int t0,
                        int k_steps, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((n_env + 64 * W - 1) / (64 * W));
  switch (d.noise) {
    case PCG_NOISE_OFF:
      ws_rollout_kernel<M, TB, PCG_NOISE_OFF, W, NSP><<<grid, 64 * W, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    case PCG_NOISE_PHILOX:
      ws_rollout_kernel<M, TB, PCG_NOISE_PHILOX, W, NSP><<<grid, 64 * W, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    default:
      ws_rollout_kernel<M, TB, PCG_NOISE_PHILOX_F32, W, NSP><<<grid, 64 * W, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
  }
  return cudaGetLastError();
}

template <class M, class TB, int W>
cudaError_t launch_ws(const pcg_desc& d0, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const pcg_desc d = fold_desc<M>(d0);
  const Derived<M, TB> dv = make_derived<M, TB>(d);
  const cudaError_t e = d.nsp == 1 ? launch_ws_s<M, TB, W, 1>(d, io, dv, n_env, t0, k_steps, s)
                                   : launch_ws_s<M, TB, W, 2>(d, io, dv, n_env, t0, k_steps, s);
  t_last_kernel = "ws_rollout_kernel (warp-specialised integrator/finisher pairs)";
  return e;
}

// persistent CTA per SM with a dynamically balanced finisher pool (rollout_wsm.cuh)
inline int sm_count() {                      // of the calling thread's current device (cached per ordinal)
  static std::atomic<int> cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = cache[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}

// eligible when every CTA gets 2..8 tiles of 64 envs (n > 9.5 k on 148 SMs) and `done` cannot depend on earlier
// violations
template <class M>
inline bool wsm_ok(const pcg_desc& d, const pcg_io& io, int64_t n_env, int64_t k_steps, int64_t* envs_per_cta,
                   int* n_tiles) {
  if (d.flags & PCG_F_DONE_ON_VIOLATION) return false;
  // the per-SM kernel's model variant may tabulate over a bounded argument range (cstr: CstrW): the parameters must fit
  if constexpr (has_wide_tab<typename M::WsmModel>::value) {
    if (!M::WsmModel::wide_ok(M::WsmModel::make(d.theta))) return false;
  }
  // the finishers address with 32-bit element indices
  const int64_t lim = int64_t{1} << 31;
  if (k_steps * PCG_MAX_OBS * n_env >= lim) return false;
  if (io.sp_st < 0 || io.sp_sj < 0 || io.sp_se < 0 ||
      (d.N + 1) * io.sp_st + PCG_MAX_NSP * io.sp_sj + (n_env + 2) * io.sp_se >= lim) return false;
  // One CTA per SM and wave: batches beyond 8 tiles per SM run as several waves of CTAs (one resident per SM, the
  // hardware starts the next as one retires), each again a 2..8-tile rollout of its own envs over all k_steps.
  const int64_t sms = sm_count();
  const int64_t waves = (n_env + sms * 64 * kWsmMaxTiles - 1) / (sms * 64 * kWsmMaxTiles);
  const int64_t ctas = sms * waves;
  const int64_t epc = 2 * ((n_env + 2 * ctas - 1) / (2 * ctas));
  const int tiles = static_cast<int>((epc + 63) / 64);
  *envs_per_cta = epc;
  *n_tiles = tiles;
  return tiles >= 2 && tiles <= kWsmMaxTiles;
}

template <class M, class TB, int NOISE, int NSP, bool CONU, bool SPS, bool ADAPT, bool GEN>
cudaError_t launch_wsm_n(const pcg_desc& d, const pcg_io& io, const Derived<M, TB>& dv, int64_t n_env, int t0,
                         int k_steps, cudaStream_t s, int64_t epc, int n_tiles) {
  using L = WsmLayout<ADAPT, GEN>;
  using Shared = WsmShared<M, NOISE, NSP, SPS, L::ring, GEN ? M::ND : 0>;
  static_assert(sizeof(Shared) <= 227 * 1024, "wsm_rollout_kernel: rings + tables exceed the 227 KB a CTA may opt in to");
  // > 48 KB of dynamic shared memory needs the opt-in, once per kernel instantiation AND device
  static std::atomic<uint64_t> opted{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const uint64_t bit = uint64_t{1} << (dev & 63);
  if (!(opted.load(std::memory_order_acquire) & bit)) {
    const cudaError_t attr = cudaFuncSetAttribute(wsm_rollout_kernel<M, TB, NOISE, NSP, CONU, SPS, ADAPT, GEN>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  static_cast<int>(sizeof(Shared)));
    if (attr != cudaSuccess) return attr;
    opted.fetch_or(bit, std::memory_order_release);
  }
  const unsigned grid = static_cast<unsigned>((n_env + epc - 1) / epc);
  wsm_rollout_kernel<M, TB, NOISE, NSP, CONU, SPS, ADAPT, GEN>
      <<<grid, 32 * L::warps, sizeof(Shared), s>>>(d, io, dv, n_env, t0, k_steps, epc, n_tiles);
  return cudaGetLastError();
}

template <class M, class TB, int NSP, bool CONU, bool SPS, bool ADAPT, bool GEN>
cudaError_t launch_wsm_s(const pcg_desc& d, const pcg_io& io, const Derived<M, TB>& dv, int64_t n_env, int t0,
                         int k_steps, cudaStream_t s, int64_t epc, int n_tiles) {
  switch (d.noise) {
    case PCG_NOISE_OFF:
      return launch_wsm_n<M, TB, PCG_NOISE_OFF, NSP, CONU, SPS, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
    case PCG_NOISE_PHILOX:
      return launch_wsm_n<M, TB, PCG_NOISE_PHILOX, NSP, CONU, SPS, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
    default:
      // launch_e keeps fp32-evaluated noise on the generic kernel for the adaptive and the general variant
      if constexpr (ADAPT || GEN) return cudaErrorNotSupported;
      else return launch_wsm_n<M, TB, PCG_NOISE_PHILOX_F32, NSP, CONU, SPS, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
  }
}

template <class M, class TB, int NSP, bool CONU, bool ADAPT, bool GEN>
cudaError_t launch_wsm_c(const pcg_desc& d, const pcg_io& io, const Derived<M, TB>& dv, int64_t n_env, int t0,
                         int k_steps, cudaStream_t s, int64_t epc, int n_tiles) {
  // the reference's setpoints are ONE schedule for every env (SP[k][t], pcgym.py:432-438): staged in shared memory
  const bool sps = io.sp_se == 0 && k_steps + 1 <= kWsmSpMax;
  return sps ? launch_wsm_s<M, TB, NSP, CONU, true, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles)
             : launch_wsm_s<M, TB, NSP, CONU, false, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
}

template <class M, class TB, bool ADAPT = false, bool GEN = false>
cudaError_t launch_wsm(const pcg_desc& d0, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s,
                       int64_t epc, int n_tiles) {
  const pcg_desc d = fold_desc<M>(d0);
  const Derived<M, TB> dv = make_derived<M, TB>(d);
  bool con_u = false;                       // does any constraint read an input column?
  for (int c = 0; c < d.n_con; ++c) con_u = con_u || d.con_mask_u[c] != 0;
  cudaError_t e;
  if (d.nsp == 1)
    e = con_u ? launch_wsm_c<M, TB, 1, true, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles)
              : launch_wsm_c<M, TB, 1, false, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
  else
    e = con_u ? launch_wsm_c<M, TB, 2, true, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles)
              : launch_wsm_c<M, TB, 2, false, ADAPT, GEN>(d, io, dv, n_env, t0, k_steps, s, epc, n_tiles);
  t_last_kernel = ADAPT ? "wsm_rollout_kernel (persistent CTA per SM, adaptive integrator warps + finisher pool)"
                  : GEN ? "wsm_rollout_kernel (persistent CTA per SM, general: disturbance rows + substeps)"
                        : "wsm_rollout_kernel (persistent CTA per SM: integrator warps + finisher pool)";
  return e;
}

// The general variant of the per-SM kernel (GEN) takes what used to send a cstr batch to the generic kernel although
// everything else fits the straight-line form: supplied model disturbances (per-env tables or the reference's shared
// schedule), observed disturbance slots, n_sub > 1 fixed substeps.  Still excluded: per-env parameters, uncertain-
// parameter tail, > 2 setpoints / constraints, constraints that read observation slots, a_delta, other rewards,
// external / fp32 noise.
template <class M>
inline bool wsm_gen_ok(const pcg_desc& d, const pcg_io& io, int64_t n_env, int64_t k_steps, int64_t* epc, int* tiles) {
  if (d.flags & PCG_F_ADAPTIVE) return false;
  if (d.nd == 0 && d.n_sub == 1) return false;                 // the plain variant's job
  if (io.theta != nullptr || d.n_tail > 0 || d.nd_obs > M::ND) return false;
  if ((d.flags & PCG_F_A_DELTA) && (d.flags & PCG_F_NORMALISE_A)) return false;
  if (d.reward != PCG_REWARD_SP || d.nsp < 1 || d.nsp > kFastNsp || d.n_con > kFastNcon) return false;
  for (int c = 0; c < d.n_con; ++c)
    if (d.con_mask_x[c] >> M::NX) return false;
  if (d.noise != PCG_NOISE_OFF && d.noise != PCG_NOISE_PHILOX) return false;
  if (!io.x_traj || !io.obs || !io.reward || !io.done || !io.violated || (d.n_con > 0 && !io.g)) return false;
  if (d.nd > 0) {
    bool any = false;
    for (int i = 0; i < M::ND; ++i) any = any || d.d_src[i] >= 0;
    if (any) {
      if (!io.dist || io.d_se < 0 || io.d_se > 1) return false;
      // per-env tables move as 16-byte pairs: even strides, aligned base
      if (io.d_se == 1 && (!al16(io.dist) || (io.d_st % 2) != 0 || (io.d_sj % 2) != 0)) return false;
    }
  }
  return wsm_ok<M>(d, io, n_env, k_steps, epc, tiles);
}

// Kernel-family request of the calling thread (pcg_set_kernel; PCG_KERNEL_AUTO = the library's own choice).  A request
// is honoured only where that family is eligible; otherwise the usual fall-back order applies, and pcg_last_kernel()
// says what actually ran.  The environment knobs PCG_WS (0 fast, 1 pair-per-CTA, 3 per-SM), PCG_WSM=0, PCG_NO_FAST and
// PCG_EPT=1 give the process-wide default (profiling aids).
inline thread_local int t_kernel_request = PCG_KERNEL_AUTO;
inline int kernel_request() {
  if (t_kernel_request != PCG_KERNEL_AUTO) return t_kernel_request;
  static const int env = [] {
    if (std::getenv("PCG_NO_FAST")) return static_cast<int>(PCG_KERNEL_GENERIC);
    switch (ws_override()) {
      case 0: return static_cast<int>(PCG_KERNEL_FAST);
      case 1: return static_cast<int>(PCG_KERNEL_WS);
      case 3: return static_cast<int>(PCG_KERNEL_WSM);
    }
    return static_cast<int>(PCG_KERNEL_AUTO);
  }();
  return env;
}

// lane-split, lane-asynchronous adaptive kernel (rollout_lanes.cuh): staged models (M::LANES), embedded explicit pair,
// lean observation vector, SP or no reward, no a_delta, Philox / external / no noise
template <class M>
inline bool lanes_ok(const pcg_desc& d, const pcg_io& io) {
  if (!(d.flags & PCG_F_ADAPTIVE) || needs_general<M>(d) || io.theta != nullptr) return false;
  if ((d.flags & PCG_F_A_DELTA) && (d.flags & PCG_F_NORMALISE_A)) return false;
  if (d.reward != PCG_REWARD_SP && d.reward != PCG_REWARD_NONE) return false;
  if (d.nsp > kFastNsp) return false;
  if (d.noise == PCG_NOISE_PHILOX_F32 || (d.noise == PCG_NOISE_EXTERNAL && !io.z)) return false;
  return M::lanes_ready(M::make(d.theta));
}

template <class M, class TB>
cudaError_t launch_lanes(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  constexpr int epb = (32 / M::LANES) * (kLanesBlock / 32);                    // envs per CTA
  const unsigned grid = static_cast<unsigned>((n_env + epb - 1) / epb);
  const Derived<M, TB> dv = make_derived<M, TB>(d);
  switch (d.noise) {
    case PCG_NOISE_OFF:
      lanes_rollout_kernel<M, TB, PCG_NOISE_OFF><<<grid, kLanesBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    case PCG_NOISE_PHILOX:
      lanes_rollout_kernel<M, TB, PCG_NOISE_PHILOX><<<grid, kLanesBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
    default:
      lanes_rollout_kernel<M, TB, PCG_NOISE_EXTERNAL><<<grid, kLanesBlock, 0, s>>>(d, io, dv, n_env, t0, k_steps);
      break;
  }
  t_last_kernel = "lanes_rollout_kernel (one lane per stage, lane-asynchronous control intervals)";
  return cudaGetLastError();
}

template <class M, class TB, bool ADAPTIVE, bool GENERAL>
cudaError_t launch_e(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const int req = kernel_request();
  const bool ept1 = ept_override() == 1 || req == PCG_KERNEL_FAST_EPT1 || req == PCG_KERNEL_GENERIC_EPT1;
  // two envs per thread share one Philox block, which needs an even global id for the first of them
  const bool two = M::EPT == 2 && !ept1 && vector_ok(io, n_env) && d.env_id0 % 2 == 0;
  const bool generic = req == PCG_KERNEL_GENERIC || req == PCG_KERNEL_GENERIC_EPT1;
  if constexpr (!ADAPTIVE && !GENERAL && M::FAST) {
    if (!generic && fast_ok<M>(d, io)) {
      if constexpr (M::EPT == 2) {
        if constexpr (M::WS > 0) {
          static const bool wsm_off = [] { const char* e = std::getenv("PCG_WSM"); return e && std::atoi(e) == 0; }();
          int64_t epc = 0;
          int tiles = 0;
          if (two && wsm_ok<M>(d, io, n_env, k_steps, &epc, &tiles) &&
              (req == PCG_KERNEL_WSM || (req == PCG_KERNEL_AUTO && !wsm_off)))
            return launch_wsm<typename M::WsmModel, TB>(d, io, n_env, t0, k_steps, s, epc, tiles);
          const bool ws = req == PCG_KERNEL_WS || ((req == PCG_KERNEL_AUTO || req == PCG_KERNEL_WSM) && ws_preferred(n_env));
          // the integrator warps run the model's WsModel (cstr: table-driven exp, 8 FP64 issue slots fewer per RHS)
          if (two && ws) return launch_ws<typename M::WsModel, TB, 1>(d, io, n_env, t0, k_steps, s);
        }
        if (two) return launch_fast<M, TB, 2>(d, io, n_env, t0, k_steps, s);
      }
      return launch_fast<M, TB, 1>(d, io, n_env, t0, k_steps, s);
    }
  }
  if constexpr (ADAPTIVE && !GENERAL && M::FAST && M::EPT == 2 && TB::EMBEDDED && !is_rosenbrock<TB>::value) {
    // the default scheme of make_env (adaptive embedded pair) in the integrator warps of the per-SM kernel
    if constexpr (M::WS > 0) {
      int64_t epc = 0;
      int tiles = 0;
      if (!generic && two && d.noise != PCG_NOISE_PHILOX_F32 && fast_ok<M>(d, io, true) &&
          wsm_ok<M>(d, io, n_env, k_steps, &epc, &tiles) && (req == PCG_KERNEL_WSM || req == PCG_KERNEL_AUTO))
        return launch_wsm<typename M::WsmModel, TB, true>(d, io, n_env, t0, k_steps, s, epc, tiles);
    }
  }
  if constexpr (ADAPTIVE && !GENERAL && has_lanes<M>::value && TB::EMBEDDED && !is_rosenbrock<TB>::value) {
    if (!generic && (req == PCG_KERNEL_AUTO || req == PCG_KERNEL_LANES) && lanes_ok<M>(d, io))
      return launch_lanes<M, TB>(d, io, n_env, t0, k_steps, s);
  }
  if constexpr (!ADAPTIVE && M::EPT == 2) {
    if (two) return launch_one<M, TB, ADAPTIVE, GENERAL, 2>(d, io, n_env, t0, k_steps, s);
  }
  return launch_one<M, TB, ADAPTIVE, GENERAL, 1>(d, io, n_env, t0, k_steps, s);
}

// the general variant's 16 instantiations live in their own translation unit per model (inst_<model>_gen.cu: parallel nvcc)
template <class M>
cudaError_t launch_wsm_general(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t, int64_t, int);
template <>
cudaError_t launch_wsm_general<Cstr>(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t, int64_t, int);   // inst_cstr_gen.cu

template <class M, class TB, bool ADAPTIVE>
cudaError_t launch_g(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  if constexpr (!ADAPTIVE && M::FAST && M::EPT == 2 && M::WS > 0 && std::is_same<TB, Tsit5>::value) {   // one tableau: 16 instantiations
    const int req = kernel_request();
    int64_t epc = 0;
    int tiles = 0;
    if ((req == PCG_KERNEL_AUTO || req == PCG_KERNEL_WSM) && ept_override() != 1 && vector_ok(io, n_env) &&
        d.env_id0 % 2 == 0 && wsm_gen_ok<M>(d, io, n_env, k_steps, &epc, &tiles))
      return launch_wsm_general<M>(d, io, n_env, t0, k_steps, s, epc, tiles);
  }
  return (needs_general<M>(d) || io.theta != nullptr) ? launch_e<M, TB, ADAPTIVE, true>(d, io, n_env, t0, k_steps, s)
                             : launch_e<M, TB, ADAPTIVE, false>(d, io, n_env, t0, k_steps, s);
}

template <class M, class = void>
struct tsit5_only : std::false_type {};
template <class M>
struct tsit5_only<M, std::enable_if_t<M::TSIT5_ONLY>> : std::true_type {};

template <class M>
cudaError_t launch_model(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const bool adaptive = (d.flags & PCG_F_ADAPTIVE) != 0;
  if constexpr (tsit5_only<M>::value) {          // abi.cu: validate() rejects the other schemes for these models
    return adaptive ? launch_g<M, Tsit5, true>(d, io, n_env, t0, k_steps, s)
                    : launch_g<M, Tsit5, false>(d, io, n_env, t0, k_steps, s);
  }
  switch (d.scheme) {
    case PCG_SCHEME_RK4:
      return launch_g<M, Rk4, false>(d, io, n_env, t0, k_steps, s);
    case PCG_SCHEME_DOPRI5:
      return adaptive ? launch_g<M, Dopri5, true>(d, io, n_env, t0, k_steps, s)
                      : launch_g<M, Dopri5, false>(d, io, n_env, t0, k_steps, s);
    case PCG_SCHEME_TSIT5:
      return adaptive ? launch_g<M, Tsit5, true>(d, io, n_env, t0, k_steps, s)
                      : launch_g<M, Tsit5, false>(d, io, n_env, t0, k_steps, s);
    case PCG_SCHEME_RODAS4:                       // always adaptive; models with the Rosenbrock interface (abi.cu validates)
      if constexpr (is_stiff_capable<M>::value) return launch_g<M, Rodas4, true>(d, io, n_env, t0, k_steps, s);
      else return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

// one translation unit per model (parallel nvcc); defined in inst_<model>.cu
cudaError_t launch_cstr(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_first_order(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_extraction(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_extraction10(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_four_tank(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_crystallization(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_biofilm(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_complex_cstr(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_nonsmooth(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_cstr_series(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_distillation(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_extraction_reactive(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_heat_exchanger(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_polymerisation(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_photo(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_invariant_batch(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_batch(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_oscillators(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_disease(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_hydraulic_tank(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_affine(int nx, int nu, const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);

}  // namespace pcg
