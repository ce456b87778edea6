// C ABI of libpcgym_b200.so: argument validation + dispatch to the per-model kernel launchers.
// Declarations and the reference interface each entry point replaces: include/pcgym_b200.h.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "launch.cuh"

namespace {

struct ModelInfo {
  int nx, nu, nd, np;
  const char* const* names;
  const double* defaults;
  pcg::launch_fn launch;
  bool stiff = false;        // provides the Rosenbrock interface (PCG_SCHEME_RODAS4)
};

// parameter names / defaults in reference dataclass field order (model_classes.py)
const char* const kCstrN[] = {"q", "V", "rho", "C", "deltaHr", "EA_over_R", "k0", "UA", "Ti", "Caf"};
const double kCstrD[] = {100, 100, 1000, 0.239, -5e4, 8750, 7.2e10, 5e4, 350, 1};
const char* const kFoN[] = {"K", "tau"};
const double kFoD[] = {1, 0.5};
const char* const kExN[] = {"Vl", "Vg", "m", "Kla", "eq_exponent", "X0", "Y6"};
const double kExD[] = {5, 5, 1, 5, 2, 0.6, 0.05};
const char* const kEx10N[] = {"Vl", "Vg", "m", "Kla", "eq_exponent", "X0", "Y11"};
const char* const kFtN[] = {"g", "gamma_1", "gamma_2", "k1", "k2", "a1", "a2", "a3", "a4", "A1", "A2", "A3", "A4"};
const double kFtD[] = {9.81, 0.2, 0.2, 0.00085, 0.00095, 0.0035, 0.0030, 0.0020, 0.0025, 1, 1, 1, 1};
const char* const kCrN[] = {"ka", "kb", "kc", "kd", "kg", "k1", "k2", "a", "b", "alfa", "ro"};
const double kCrD[] = {0.923714966, -6754.878558, 0.92229965554, 1.341205945, 48.07514464, -4921.261419,
                       1.871281405, 0.50523693, 7.271241375, 7.510905767, 2.658};
const char* const kBfN[] = {"V", "Va", "Kla", "m", "eq_exponent", "O_air", "vm_1", "vm_2", "K1", "K2", "KO_1", "KO_2"};
const double kBfD[] = {10.0, 15.0, 1.5, 0.5, 1.0, 300, 0.8, 1.0, 0.5, 0.1, 1.5, 0.5};

// further registry models (models_ext.cuh)
const char* const kCcN[] = {"q", "V", "rho", "C", "deltaHr1", "EA1_over_R", "k01", "deltaHr2", "EA2_over_R", "k02", "UA", "Ti", "Caf"};
const double kCcD[] = {100, 100, 1000, 0.239, -5e4, 8750, 7.2e10, -3e4, 9000, 1.0e10, 5e4, 350, 1};
const char* const kNsN[] = {"a_11", "a_12", "a_21", "a_22", "b_1", "b_2"};
const double kNsD[] = {0, 1, -2, -3, 0, 1};
const char* const kCsN[] = {"C_O", "T_O", "V1", "V2", "U1A1", "U2A2", "rho", "cp", "k", "E", "deltaH", "R"};
const double kCsD[] = {97.35, 298, 1e-3, 2e-3, 0.461, 0.732, 1.05e3, 3.766, 3.118e5, 46.14, 58.41, 8.3145e-3};
const char* const kDcN[] = {"D", "q", "alpha", "X_feed", "M0", "Mb", "M"};
const double kDcD[] = {100.0, 1.0, 5.0, 0.2, 2000.0, 2000.0, 2000.0};
const char* const kErN[] = {"Vl", "Vg", "m", "Kla", "k", "eq_exponent", "XA0", "YA6", "YB6", "YC6"};
const double kErD[] = {5.0, 5.0, 1.0, 0.01, 0.1, 2.0, 2.00, 0.00, 2.00, 0.00};
const char* const kHxN[] = {"Utm", "Usm", "L", "Dt", "Dm", "Ds", "cpt", "cpm", "cps", "rhot", "rhom", "rhos"};
const double kHxD[] = {1.0, 1.0, 1.0, 1.0, 2.0, 3.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
const char* const kPoN[] = {"Ap", "Ad", "At", "Ep_over_R", "Ed_over_R", "Et_over_R", "f", "V", "deltaHp", "rho", "cp"};
const double kPoD[] = {6e10, 4e10, 9e10, 7750, 8500, 8250, 0.5, 1.0, -3e4, 1200.0, 2.0};
const char* const kPhN[] = {"u_m", "u_d", "Y_NX", "k_m", "k_d", "k_sq", "K_Nq", "k_iq", "k_s", "k_i", "k_N"};
const double kPhD[] = {0.0572, 0.0, 504.5, 0.00016, 0.281, 23.51, 16.89, 800.0, 178.9, 447.1, 393.1};
const char* const kIbN[] = {"k1f", "k1r", "k2f", "k2r"};
const double kIbD[] = {55.0, 1.0, 2.0, 1.0};
const char* const kBaN[] = {"k01", "k02", "EA1", "EA2", "R", "dH1", "dH2", "rho", "Cp", "UA", "V"};
const double kBaD[] = {1.0, 0.5, 5000, 6000, 8.314, -1000, -1500, 1000, 4.0, 100, 1.0};
const char* const kOsN[] = {"N", "k", "m"};
const double kOsD[] = {10, 1.0, 1.0};
const char* const kDiN[] = {"beta", "gamma"};
const double kDiD[] = {0.3, 0.1};
const char* const kHtN[] = {"D"};
const double kHtD[] = {1.0};

// affine custom models: theta = [A row-major | B row-major | c], defaults 0; parameter names a_i_j, b_i_j, c_i
struct AffineNames {
  char text[PCG_MAX_PARAM][8];
  const char* ptr[PCG_MAX_PARAM];
  AffineNames(int nx, int nu) {
    int k = 0;
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < nx; ++j) std::snprintf(text[k++], 8, "a_%d_%d", i, j);
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < nu; ++j) std::snprintf(text[k++], 8, "b_%d_%d", i, j);
    for (int i = 0; i < nx; ++i) std::snprintf(text[k++], 8, "c_%d", i);
    for (int i = 0; i < PCG_MAX_PARAM; ++i) ptr[i] = i < k ? text[i] : "";
  }
};
const double kZeros[PCG_MAX_PARAM] = {};
template <int NX, int NU>
cudaError_t launch_affine_t(const pcg_desc& d, const pcg_io& io, int64_t n, int t0, int k, cudaStream_t s) {
  return pcg::launch_affine(NX, NU, d, io, n, t0, k, s);
}
template <int NX, int NU>
ModelInfo affine_info() {
  static const AffineNames names(NX, NU);
  return ModelInfo{NX, NU, 0, NX * NX + NX * NU + NX, names.ptr, kZeros, launch_affine_t<NX, NU>};
}
#define PCG_AFFINE_ROW(NX, NU) affine_info<NX, NU>()

template <class M>
constexpr ModelInfo info(const char* const* n, const double* dflt, pcg::launch_fn f) {
  return ModelInfo{M::NX, M::NU, M::ND, M::NP, n, dflt, f, pcg::is_stiff_capable<M>::value};
}

const ModelInfo kModels[PCG_MODEL_COUNT] = {
    info<pcg::Cstr>(kCstrN, kCstrD, pcg::launch_cstr),
    info<pcg::FirstOrder>(kFoN, kFoD, pcg::launch_first_order),
    info<pcg::Extraction>(kExN, kExD, pcg::launch_extraction),
    info<pcg::FourTank>(kFtN, kFtD, pcg::launch_four_tank),
    info<pcg::Crystallization>(kCrN, kCrD, pcg::launch_crystallization),
    info<pcg::Biofilm>(kBfN, kBfD, pcg::launch_biofilm),
    info<pcg::ComplexCstr>(kCcN, kCcD, pcg::launch_complex_cstr),
    info<pcg::Nonsmooth>(kNsN, kNsD, pcg::launch_nonsmooth),
    info<pcg::CstrSeries>(kCsN, kCsD, pcg::launch_cstr_series),
    info<pcg::Distillation>(kDcN, kDcD, pcg::launch_distillation),
    info<pcg::ExtractionReactive>(kErN, kErD, pcg::launch_extraction_reactive),
    info<pcg::HeatExchanger>(kHxN, kHxD, pcg::launch_heat_exchanger),
    info<pcg::Polymerisation>(kPoN, kPoD, pcg::launch_polymerisation),
    info<pcg::PhotoProduction>(kPhN, kPhD, pcg::launch_photo),
    info<pcg::InvariantBatch>(kIbN, kIbD, pcg::launch_invariant_batch),
    info<pcg::BatchReactor>(kBaN, kBaD, pcg::launch_batch),
    info<pcg::Oscillators>(kOsN, kOsD, pcg::launch_oscillators),
    info<pcg::Disease>(kDiN, kDiD, pcg::launch_disease),
    info<pcg::HydraulicTank>(kHtN, kHtD, pcg::launch_hydraulic_tank),
    PCG_AFFINE_ROW(1, 1), PCG_AFFINE_ROW(1, 2), PCG_AFFINE_ROW(2, 1), PCG_AFFINE_ROW(2, 2),
    PCG_AFFINE_ROW(3, 1), PCG_AFFINE_ROW(3, 2), PCG_AFFINE_ROW(4, 1), PCG_AFFINE_ROW(4, 2),
    info<pcg::Extraction10>(kEx10N, kExD, pcg::launch_extraction10),       // extension, after the affine block (ids are ABI)
};

// the measurement-noise transform on caller-supplied Philox words (pcg_normal_pairs)
__global__ void normal_pairs_kernel(const uint32_t* w0, const uint32_t* w1, double* z0, double* z1, int64_t n, bool f32) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a, b;
  if (f32) pcg::normal_pair_u32<true>(w0[i], w1[i], pcg::kIcdfTab, a, b);
  else pcg::normal_pair_u32<false>(w0[i], w1[i], pcg::kIcdfTab, a, b);
  z0[i] = a;
  z1[i] = b;
}

// the branch-free fp64 helpers of fastmath.cuh on caller-supplied arguments (pcg_math_probe)
__global__ void math_probe_kernel(int op, const double* x, double* y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (op) {
    case PCG_MATH_EXP: r = pcg::exp_sat(v); break;
    case PCG_MATH_EXP_HORNER: r = pcg::exp_sat_horner(v); break;
    case PCG_MATH_EXP_TAB: { double sc, po; pcg::exp_tab_split(v, pcg::kExp2Tab, sc, po); r = sc * po; break; }
    case PCG_MATH_LOG: r = pcg::log_tab(v); break;
    case PCG_MATH_SQRT: r = pcg::sqrt_fast(v); break;
    case PCG_MATH_RCP: r = pcg::rcp_nr(v); break;
    default: r = pcg::rcp_o3(v); break;
  }
  y[i] = r;
}

std::atomic<int64_t> g_launches{0};
thread_local char t_cuda_err[256] = "";

bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

int validate(const pcg_desc* d, const pcg_io* io, int64_t n_env, int64_t t0, int64_t k_steps) {
  if (!d || !io) return PCG_E_NULL;
  if (d->abi_version != PCG_ABI_VERSION || d->struct_bytes != static_cast<int32_t>(sizeof(pcg_desc)))
    return PCG_E_VERSION;
  if (d->model < 0 || d->model >= PCG_MODEL_COUNT) return PCG_E_MODEL;
  if (d->scheme < 0 || d->scheme >= PCG_SCHEME_COUNT) return PCG_E_SCHEME;
  const bool adaptive = (d->flags & PCG_F_ADAPTIVE) != 0;
  if (adaptive && d->scheme == PCG_SCHEME_RK4) return PCG_E_SCHEME;
  if (d->scheme == PCG_SCHEME_RODAS4 && (!adaptive || !kModels[d->model].stiff)) return PCG_E_SCHEME;   // adaptive, Jacobian needed
  if (d->model >= PCG_MODEL_AFFINE_BASE && d->model < PCG_MODEL_EXTRACTION_10 && d->scheme != PCG_SCHEME_TSIT5)
    return PCG_E_SCHEME;                                       // the affine custom models have Tsit5 kernels only
  const ModelInfo& m = kModels[d->model];
  if (d->nu_a != m.nu) return PCG_E_DIMS;
  if (d->model == PCG_MODEL_OSCILLATORS && d->theta[0] != pcg::Oscillators::NOSC) return PCG_E_DIMS;   // state count is 2 N
  if (d->nd != 0 && d->nd != m.nd) return PCG_E_DIMS;
  if (d->nsp < 0 || d->nsp > PCG_MAX_NSP || d->nd_obs < 0 || d->nd_obs > m.nd || d->n_tail < 0 ||
      d->n_tail > PCG_MAX_NTAIL || d->n_con < 0 || d->n_con > PCG_MAX_NCON)
    return PCG_E_DIMS;
  if (d->nd == 0 && d->nd_obs != 0) return PCG_E_DIMS;
  for (int j = 0; j < d->nsp; ++j)
    if (d->sp_state[j] < 0 || d->sp_state[j] >= m.nx) return PCG_E_DIMS;
  for (int j = 0; j < d->nd_obs; ++j)
    if (d->d_obs[j] < 0 || d->d_obs[j] >= m.nd) return PCG_E_DIMS;
  if (adaptive ? (d->max_steps < 1 || !(d->rtol > 0.0) || !(d->atol >= 0.0)) : (d->n_sub < 1)) return PCG_E_RANGE;
  if (!(d->dt > 0.0) || d->N < 2) return PCG_E_RANGE;
  if (n_env < 1 || n_env > (int64_t{1} << 31) - pcg::kBlock || k_steps < 1 || k_steps > (1 << 24)) return PCG_E_RANGE;
  // the kernels read SP / disturbance row t + 1 and write cons_info row t + 1 of every interval: t0 + k_steps <= N - 1
  if (t0 < 0 || t0 + k_steps > static_cast<int64_t>(d->N) - 1) return PCG_E_RANGE;
  if (!io->x || !io->a) return PCG_E_NULL;
  if (d->nsp > 0 && !io->sp) return PCG_E_NULL;
  if (d->nd > 0) {
    bool any = false;
    for (int i = 0; i < m.nd; ++i) any = any || d->d_src[i] >= 0;
    if (any && !io->dist) return PCG_E_NULL;
  }
  if (d->noise == PCG_NOISE_EXTERNAL && !io->z) return PCG_E_NULL;
  if (d->noise < 0 || d->noise > PCG_NOISE_PHILOX_F32) return PCG_E_RANGE;
  if (d->n_tail > 0 && !io->tail) return PCG_E_NULL;
  if ((d->flags & PCG_F_A_DELTA) && (d->flags & PCG_F_NORMALISE_A) && !io->a_save) return PCG_E_NULL;
  if (d->reward < 0 || d->reward > PCG_REWARD_OCP) return PCG_E_RANGE;
  if (d->reward == PCG_REWARD_OCP && (d->ocp_mode & 3) != 0 && !io->u_prev) return PCG_E_NULL;
  const void* ptrs[] = {io->x, io->a, io->a_save, io->z, io->tail, io->x_traj, io->obs, io->reward, io->g, io->g0,
                        io->h_carry, io->u_prev, io->theta};
  for (const void* p : ptrs)
    if (p && !aligned8(p)) return PCG_E_ALIGN;
  return PCG_OK;
}

}  // namespace

extern "C" {

int pcg_abi_version(void) { return PCG_ABI_VERSION; }
int pcg_sizeof_desc(void) { return static_cast<int>(sizeof(pcg_desc)); }
int pcg_sizeof_io(void) { return static_cast<int>(sizeof(pcg_io)); }

int pcg_model_dims(int model, int* nx, int* nu, int* nd, int* nparam) {
  if (model < 0 || model >= PCG_MODEL_COUNT) return PCG_E_MODEL;
  const ModelInfo& m = kModels[model];
  if (nx) *nx = m.nx;
  if (nu) *nu = m.nu;
  if (nd) *nd = m.nd;
  if (nparam) *nparam = m.np;
  return PCG_OK;
}

int pcg_model_stiff_capable(int model) {
  return (model >= 0 && model < PCG_MODEL_COUNT && kModels[model].stiff) ? 1 : 0;
}

const char* pcg_model_param_name(int model, int i) {
  if (model < 0 || model >= PCG_MODEL_COUNT || i < 0 || i >= kModels[model].np) return nullptr;
  return kModels[model].names[i];
}

int pcg_model_param_default(int model, int i, double* value) {
  if (model < 0 || model >= PCG_MODEL_COUNT) return PCG_E_MODEL;
  if (i < 0 || i >= kModels[model].np) return PCG_E_RANGE;
  if (!value) return PCG_E_NULL;
  *value = kModels[model].defaults[i];
  return PCG_OK;
}

int pcg_rollout(const pcg_desc* d, const pcg_io* io, int64_t n_env, int64_t t0, int64_t k_steps, void* cuda_stream) {
  const int rc = validate(d, io, n_env, t0, k_steps);
  if (rc != PCG_OK) return rc;
  const cudaError_t e = kModels[d->model].launch(*d, *io, n_env, static_cast<int>(t0), static_cast<int>(k_steps),
                                                 static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) {
    std::strncpy(t_cuda_err, cudaGetErrorString(e), sizeof(t_cuda_err) - 1);
    return PCG_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCG_OK;
}

int pcg_step(const pcg_desc* d, const pcg_io* io, int64_t n_env, int64_t t, void* cuda_stream) {
  return pcg_rollout(d, io, n_env, t, 1, cuda_stream);
}

int pcg_normal_pairs(const uint32_t* w0, const uint32_t* w1, double* z0, double* z1, int64_t n, int f32,
                     void* cuda_stream) {
  if (!w0 || !w1 || !z0 || !z1) return PCG_E_NULL;
  if (n < 1 || n > (int64_t{1} << 31) - 256) return PCG_E_RANGE;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  normal_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(w0, w1, z0, z1, n, f32 != 0);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::strncpy(t_cuda_err, cudaGetErrorString(e), sizeof(t_cuda_err) - 1);
    return PCG_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCG_OK;
}

int pcg_math_probe(int op, const double* x, double* y, int64_t n, void* cuda_stream) {
  if (!x || !y) return PCG_E_NULL;
  if (op < 0 || op >= PCG_MATH_COUNT || n < 1 || n > (int64_t{1} << 31) - 256) return PCG_E_RANGE;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  math_probe_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(op, x, y, n);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::strncpy(t_cuda_err, cudaGetErrorString(e), sizeof(t_cuda_err) - 1);
    return PCG_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCG_OK;
}

int64_t pcg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* pcg_strerror(int status) {
  switch (status) {
    case PCG_OK: return "ok";
    case PCG_E_NULL: return "required pointer is NULL";
    case PCG_E_VERSION: return "pcg_desc ABI version / size mismatch";
    case PCG_E_MODEL: return "unknown model id";
    case PCG_E_SCHEME: return "unknown integrator scheme (or adaptive RK4)";
    case PCG_E_DIMS: return "descriptor dimensions inconsistent with the model";
    case PCG_E_RANGE: return "n_env / k_steps / t / integrator option out of range";
    case PCG_E_ALIGN: return "fp64 device pointer not 8-byte aligned";
    case PCG_E_CUDA: return "CUDA runtime error";
  }
  return "unknown status";
}

const char* pcg_last_cuda_error(void) { return t_cuda_err; }

const char* pcg_last_kernel(void) { return pcg::t_last_kernel; }

int pcg_set_kernel(int kernel) {
  if (kernel < PCG_KERNEL_AUTO || kernel > PCG_KERNEL_LANES) return PCG_E_RANGE;
  const int prev = pcg::t_kernel_request;
  pcg::t_kernel_request = kernel;
  return prev;
}

}  // extern "C"
