// rollout_kernel instantiations for the affine custom models (models_ext.cuh: Affine<NX, NU>), Tsit5 only.
#include "launch.cuh"

namespace pcg {
cudaError_t launch_affine(int nx, int nu, const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps,
                          cudaStream_t s) {
#define PCG_AFF(NX, NU) \
  if (nx == NX && nu == NU) return launch_model<Affine<NX, NU>>(d, io, n_env, t0, k_steps, s);
  PCG_AFF(1, 1) PCG_AFF(1, 2) PCG_AFF(2, 1) PCG_AFF(2, 2) PCG_AFF(3, 1) PCG_AFF(3, 2) PCG_AFF(4, 1) PCG_AFF(4, 2)
#undef PCG_AFF
  return cudaErrorInvalidValue;
}
}  // namespace pcg
