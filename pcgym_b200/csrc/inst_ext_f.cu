// rollout_kernel instantiations for further registry models (models_ext.cuh).
#include "launch.cuh"

namespace pcg {
cudaError_t launch_oscillators(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<Oscillators>(d, io, n_env, t0, k_steps, s);
}
}  // namespace pcg
