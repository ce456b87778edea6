// rollout_kernel instantiations for the "first_order" model.
#include "launch.cuh"

namespace pcg {
cudaError_t launch_first_order(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<FirstOrder>(d, io, n_env, t0, k_steps, s);
}
}  // namespace pcg
