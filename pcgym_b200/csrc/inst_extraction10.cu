// rollout_kernel instantiations for the ten-stage extension of the "extraction" model.
#include "launch.cuh"

namespace pcg {
cudaError_t launch_extraction10(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s) {
  return launch_model<Extraction10>(d, io, n_env, t0, k_steps, s);
}
}  // namespace pcg
