// rollout_kernel instantiations for further registry models (models_ext.cuh).
#include "launch.cuh"

namespace pcg {
cudaError_t launch_nonsmooth(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<Nonsmooth>(d, io, n_env, t0, k_steps, s);
}
cudaError_t launch_invariant_batch(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<InvariantBatch>(d, io, n_env, t0, k_steps, s);
}
cudaError_t launch_disease(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<Disease>(d, io, n_env, t0, k_steps, s);
}
cudaError_t launch_hydraulic_tank(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<HydraulicTank>(d, io, n_env, t0, k_steps, s);
}
}  // namespace pcg
