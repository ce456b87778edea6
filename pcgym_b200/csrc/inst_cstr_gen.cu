// General variant of the per-SM kernel for the "cstr" model (disturbance rows through the ring, substeps): its own
// translation unit so that it compiles beside inst_cstr.cu.
#include "launch.cuh"

namespace pcg {
template <>
cudaError_t launch_wsm_general<Cstr>(const pcg_desc& d, const pcg_io& io, int64_t n_env, int t0, int k_steps, cudaStream_t s,
                                     int64_t epc, int tiles) {
  return launch_wsm<Cstr::WsmModel, Tsit5, false, true>(d, io, n_env, t0, k_steps, s, epc, tiles);
}
}  // namespace pcg
