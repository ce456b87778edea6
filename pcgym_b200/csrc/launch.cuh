// Host-side launch of rollout_kernel for one model: picks the (scheme, adaptive) instantiation.
#pragma once
#include <cuda_runtime.h>

#include "rollout_kernel.cuh"

namespace pcg {

using launch_fn = cudaError_t (*)(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);

template <class M, class TB, bool ADAPTIVE>
cudaError_t launch_one(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((n_env + kBlock - 1) / kBlock);
  rollout_kernel<M, TB, ADAPTIVE><<<grid, kBlock, 0, s>>>(d, io, n_env, t0, k_steps);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_model(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  const bool adaptive = (d.flags & PCG_F_ADAPTIVE) != 0;
  switch (d.scheme) {
    case PCG_SCHEME_RK4:
      return launch_one<M, Rk4, false>(d, io, n_env, t0, k_steps, s);
    case PCG_SCHEME_DOPRI5:
      return adaptive ? launch_one<M, Dopri5, true>(d, io, n_env, t0, k_steps, s)
                      : launch_one<M, Dopri5, false>(d, io, n_env, t0, k_steps, s);
    case PCG_SCHEME_TSIT5:
      return adaptive ? launch_one<M, Tsit5, true>(d, io, n_env, t0, k_steps, s)
                      : launch_one<M, Tsit5, false>(d, io, n_env, t0, k_steps, s);
  }
  return cudaErrorInvalidValue;
}

// one translation unit per model (parallel nvcc); defined in inst_<model>.cu
cudaError_t launch_cstr(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_first_order(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_extraction(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_four_tank(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_crystallization(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);
cudaError_t launch_biofilm(const pcg_desc&, const pcg_io&, int64_t, int, int, cudaStream_t);

}  // namespace pcg
