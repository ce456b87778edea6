// rollout_kernel instantiations for further registry models (models_ext.cuh).
#include "launch.cuh"

namespace pcg {
cudaError_t launch_complex_cstr(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<ComplexCstr>(d, io, n_env, t0, k_steps, s);
}
cudaError_t launch_batch(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<BatchReactor>(d, io, n_env, t0, k_steps, s);
}
cudaError_t launch_cstr_series(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<CstrSeries>(d, io, n_env, t0, k_steps, s);
}
}  // namespace pcg
