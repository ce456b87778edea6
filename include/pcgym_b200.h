/* pcgym_b200.h -- C ABI of the B200 batched process-control rollout engine.
 *
 * Drop-in boundary for ONE hot path of MaximilianB2/pc-gym: `make_env.step()`
 * (reference src/pcgym/pcgym.py:350-500) -> `integration_engine.casadi_step /
 * jax_step` (src/pcgym/integrator.py:65-107) -> model RHS
 * (src/pcgym/model_classes.py), advanced for N independent environments per
 * launch.  The reference binds this path in-process from Python, so the binding a
 * maintainer adds is a ctypes stub (shown in INTEGRATION.md).
 *
 * Conventions
 *   - plain C, no torch types; the caller owns every buffer (device memory);
 *   - every entry point returns 0 or a negative pcg_status synchronously for
 *     argument / launch errors; it never throws, prints or exits;
 *   - work is enqueued on `cuda_stream` (a cudaStream_t, NULL = legacy default
 *     stream) and the call returns without synchronising;
 *   - numerical trouble is reported per environment in `io.status`
 *     (bit0 non-finite state, bit1 adaptive step cap hit, bit2 out of a kernel's table range), never as a host error;
 *   - no global mutable state: re-entrant across streams and threads; no
 *     allocation on the hot path.
 *   - data layout: struct-of-arrays, fp64, environment index contiguous
 *     ("[field][env]"); fp64 pointers must be 8-byte aligned (allocate 16-byte aligned and keep
 *     n_env even so every row is eligible for 128-bit access).
 */
#ifndef PCGYM_B200_H
#define PCGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCG_ABI_VERSION 2

#define PCG_MAX_NX    24   /* model states (heat_exchanger)                     */
#define PCG_MAX_NU     8   /* manipulated inputs + model disturbance channels   */
#define PCG_MAX_NSP    4   /* setpoint keys                                     */
#define PCG_MAX_ND     4   /* model disturbance channels                        */
#define PCG_MAX_NTAIL  8   /* observation-only constants (uncertain parameters) */
#define PCG_MAX_NCON   8   /* affine constraints                                */
#define PCG_MAX_PARAM 32   /* model parameters theta (affine custom model 4x2: 28) */
#define PCG_AFFINE_MAX_NX 4
#define PCG_AFFINE_MAX_NU 2
#define PCG_MAX_OBS   (PCG_MAX_NX + PCG_MAX_NSP + PCG_MAX_ND + PCG_MAX_NTAIL)

/* model registry: reference src/pcgym/pcgym.py:128-148 (keys in comments) */
enum pcg_model {
  PCG_MODEL_CSTR = 0,             /* "cstr"                   model_classes.py:23   */
  PCG_MODEL_FIRST_ORDER = 1,      /* "first_order_system"     model_classes.py:296  */
  PCG_MODEL_EXTRACTION = 2,       /* "multistage_extraction"  model_classes.py:346  */
  PCG_MODEL_FOUR_TANK = 3,        /* "four_tank"              model_classes.py:864  */
  PCG_MODEL_CRYSTALLIZATION = 4,  /* "crystallization"        model_classes.py:1232 */
  PCG_MODEL_BIOFILM = 5,          /* "biofilm_reactor"        model_classes.py:1046 */
  PCG_MODEL_COMPLEX_CSTR = 6,     /* "complex_cstr"           model_classes.py:65   */
  PCG_MODEL_NONSMOOTH = 7,        /* "nonsmooth_control"      model_classes.py:509  */
  PCG_MODEL_CSTR_SERIES = 8,      /* "cstr_series_recycle"    model_classes.py:611  */
  PCG_MODEL_DISTILLATION = 9,     /* "distillation_column"    model_classes.py:682  */
  PCG_MODEL_EXTRACTION_REACTIVE = 10, /* "multistage_extraction_reactive" model_classes.py:763 */
  PCG_MODEL_HEAT_EXCHANGER = 11,  /* "heat_exchanger"         model_classes.py:935  */
  PCG_MODEL_POLYMERISATION = 12,  /* "polymerisation_reactor" model_classes.py:1158 */
  PCG_MODEL_PHOTO = 13,           /* "photobioreactor"        model_classes.py:433  */
  PCG_MODEL_INVARIANT_BATCH = 14, /* "invariant_batch"        model_classes.py:268 (no inputs: 1 ignored input row) */
  PCG_MODEL_BATCH = 15,           /* "batch"                  model_classes.py:222  */
  PCG_MODEL_OSCILLATORS = 16,     /* "coupled_oscillator"     model_classes.py:186 (N = 10; no inputs: 1 ignored row) */
  PCG_MODEL_DISEASE = 17,         /* "disease"                model_classes.py:156  */
  PCG_MODEL_HYDRAULIC_TANK = 18,  /* "hydraulic_tank"         model_classes.py:128  */
  /* env_params["custom_model"] (pcgym.py:150-153): a Python RHS cannot run in a kernel, but an AFFINE one
   * dx = A x + B u + c -- the reference's own example, tests/environment/test_make_env_custom_model.py:7-26 -- is
   * probed on the host and run by these; id = BASE + (nx - 1) * PCG_AFFINE_MAX_NU + (nu - 1);
   * theta = A row-major (nx*nx), B row-major (nx*nu), c (nx); Tsit5 only. */
  PCG_MODEL_AFFINE_BASE = 19,
  /* EXTENSION, not a reference model: ten stages of multistage_extraction's stage balances (20 states), for
   * BASELINE.json config 4's wording "10-stage, ~20-state"; the reference column has five (model_classes.py:381,424) */
  PCG_MODEL_EXTRACTION_10 = 27,   /* "multistage_extraction_10stage" */
  PCG_MODEL_COUNT = 28
};

/* integrator: replaces integrator.py:90-107 (CVODES) / :65-88 (diffrax Tsit5 + PID 1e-8) */
enum pcg_scheme {
  PCG_SCHEME_RK4 = 0,     /* classical RK4, fixed step only                         */
  PCG_SCHEME_DOPRI5 = 1,  /* Dormand-Prince 5(4) ("RK45")                           */
  PCG_SCHEME_TSIT5 = 2,   /* Tsitouras 5(4), the pair the reference's jax path uses */
  PCG_SCHEME_RODAS4 = 3,  /* RODAS4 Rosenbrock 4(3), L-stable, analytic Jacobian: the stiff counterpart of the
                             reference's CVODES BDF (integrator.py:163-182); adaptive only; models that provide
                             the Jacobian (pcg_model_stiff_capable)                  */
  PCG_SCHEME_COUNT = 4
};

enum pcg_noise {
  PCG_NOISE_OFF = 0,
  PCG_NOISE_PHILOX = 1,      /* Philox4x32-10 keyed (seed; env, t, lane), fp64 inverse normal CDF */
  PCG_NOISE_EXTERNAL = 2,    /* standard normals supplied in io.z (parity tests)          */
  PCG_NOISE_PHILOX_F32 = 3   /* same counter scheme, Box-Muller evaluated in fp32         */
};

enum pcg_reward {
  PCG_REWARD_NONE = 0,   /* custom reward handled by the caller              */
  PCG_REWARD_SP = 1,     /* SP_reward_fn    pcgym.py:535-558                 */
  PCG_REWARD_BATCH = 2,  /* batch_reward_fn pcgym.py:502-532                 */
  PCG_REWARD_OCP = 3     /* the paper scripts' custom reward, fused: normalised squared setpoint error
                            on the NOISY observation + R * (delta u_norm)^2
                            (pc-gym_paper/train_policies/cstr/custom_reward.py:3-39)          */
};

/* pcg_desc.ocp_mode bit: the crystallization script's reward (pc-gym_paper/train_policies/crystalisation/
 * cryst_train.py:17-48): setpoint 0 tracks CV = sqrt(mu2 mu0 / mu1^2 - 1), setpoint 1 tracks Ln = mu1 / mu0, both
 * computed from the NOISY observation of states 0..2 instead of reading the CV / Ln states */
#define PCG_OCP_MOMENTS 4

enum pcg_flags {
  PCG_F_NORMALISE_A = 1 << 0,   /* pcgym.py:372-375                                 */
  PCG_F_NORMALISE_O = 1 << 1,   /* pcgym.py:483-493                                 */
  PCG_F_A_DELTA = 1 << 2,       /* pcgym.py:376-383 (only live with NORMALISE_A)    */
  PCG_F_R_PENALTY = 1 << 3,     /* pcgym.py:556-557 / :529-530                      */
  PCG_F_DONE_ON_VIOLATION = 1 << 4, /* pcgym.py:613-614                             */
  PCG_F_ADAPTIVE = 1 << 5,      /* embedded error control (DOPRI5 / TSIT5 only)     */
  PCG_F_EPISODE_START = 1 << 7, /* this launch begins an episode: io.done_state is NOT read (every env enters not
                                   done), only written -- saves the caller a memset between back-to-back episodes  */
  PCG_F_SP_HIDDEN = 1 << 6      /* x0 / o_space carry no setpoint slots: the observation is [x] only while the
                                   reward still tracks SP (the reference's own custom-model test runs like this,
                                   tests/environment/test_make_env_custom_model.py:66-86; numpy lets the
                                   setpoint write at pcgym.py:437 fall off the end of the state vector)          */
};

enum pcg_status {
  PCG_OK = 0,
  PCG_E_NULL = -1,        /* required pointer missing                   */
  PCG_E_VERSION = -2,     /* desc.abi_version / struct size mismatch    */
  PCG_E_MODEL = -3,       /* unknown model id                           */
  PCG_E_SCHEME = -4,      /* unknown scheme / adaptive RK4              */
  PCG_E_DIMS = -5,        /* dimension fields inconsistent with model   */
  PCG_E_RANGE = -6,       /* n_env, k_steps, t out of range             */
  PCG_E_ALIGN = -7,       /* fp64 pointer not 8-byte aligned            */
  PCG_E_CUDA = -8         /* CUDA runtime error (pcg_last_cuda_error)   */
};

/* kernel families (pcg_set_kernel / pcg_last_kernel): csrc/launch.cuh picks one per configuration and batch size */
enum pcg_kernel {
  PCG_KERNEL_AUTO = -1,         /* the library's own choice (default)                                            */
  PCG_KERNEL_GENERIC = 0,       /* rollout_kernel: every option, run-time flags                                  */
  PCG_KERNEL_FAST = 1,          /* fast_rollout_kernel: straight-line, software-pipelined                        */
  PCG_KERNEL_WS = 2,            /* ws_rollout_kernel: integrator / finisher warp pairs                           */
  PCG_KERNEL_WSM = 3,           /* wsm_rollout_kernel: persistent CTA per SM, finisher pool                      */
  PCG_KERNEL_FAST_EPT1 = 4,     /* fast_rollout_kernel with one environment per thread (64-bit row accesses)    */
  PCG_KERNEL_GENERIC_EPT1 = 5,  /* rollout_kernel with one environment per thread                               */
  PCG_KERNEL_LANES = 6          /* lanes_rollout_kernel: one lane per model stage, lane-asynchronous intervals
                                   (adaptive pairs on multistage_extraction and its ten-stage extension)        */
};

/* per-env status bits written to io.status */
#define PCG_ST_NONFINITE 1
#define PCG_ST_STEPCAP   2
#define PCG_ST_RANGE     4   /* the state left the band a table-driven kernel covers (build option CstrW of the per-SM kernel: reactor
                                temperature outside 250..760 K at the end of an interval); the env's results are not
                                valid from there on -- such dynamics need scheme RODAS4 or pcg_set_kernel(PCG_KERNEL_GENERIC) */

/* Host-side descriptor, passed by value to the kernel (constant bank).  Everything
 * the reference keeps on the make_env object that the step needs. */
typedef struct pcg_desc {
  int32_t abi_version;        /* PCG_ABI_VERSION */
  int32_t struct_bytes;       /* sizeof(pcg_desc) as seen by the caller */
  int32_t model;              /* enum pcg_model */
  int32_t scheme;             /* enum pcg_scheme */
  int32_t flags;              /* enum pcg_flags bits */
  int32_t n_sub;              /* fixed mode: substeps per control interval (>=1) */
  int32_t max_steps;          /* adaptive mode: cap on attempted steps per interval */
  int32_t N;                  /* episode horizon env_params["N"]; done at t == N-1 (pcgym.py:448) */
  double dt;                  /* control interval tsim / N (pcgym.py:110) */
  double rtol, atol;          /* adaptive mode tolerances */
  double theta[PCG_MAX_PARAM];/* model parameters in pcg_model_param_names() order */

  /* inputs: uk = [actions(nu_a) | model disturbances(nd)]   pcgym.py:386-414 */
  int32_t nu_a;               /* manipulated inputs supplied in io.a */
  int32_t nd;                 /* 0, or the model's disturbance count when disturbances are active */
  int32_t d_src[PCG_MAX_ND];  /* per model disturbance: row of io.dist, or -1 = d_default */
  double d_default[PCG_MAX_ND];
  double a_scale[PCG_MAX_NU]; /* (hi-lo)/2          */
  double a_off[PCG_MAX_NU];   /* lo + (hi-lo)/2     */
  double a_act_lo[PCG_MAX_NU];/* a_space_act (a_delta clip) */
  double a_act_hi[PCG_MAX_NU];

  /* observation vector: [x(nx) | SP(nsp) | supplied disturbances(nd_obs) | tail(n_tail)] */
  int32_t nsp;
  int32_t sp_state[PCG_MAX_NSP];   /* state index each setpoint tracks */
  double r_scale[PCG_MAX_NSP];
  double sp_init[PCG_MAX_NSP];     /* SP slots of x0: only seen by the t==0 constraint pre-check */
  int32_t nd_obs;
  int32_t d_obs[PCG_MAX_ND];       /* which model disturbance fills each observed slot */
  int32_t n_tail;
  uint32_t obs_zero_mask;          /* bit i set: observation i is zeroed (partial_observation) */
  double o_scale[PCG_MAX_OBS];     /* 2/(hi-lo)           */
  double o_off[PCG_MAX_OBS];       /* -2 lo/(hi-lo) - 1   */

  /* measurement noise pcgym.py:453-466 */
  int32_t noise;                   /* enum pcg_noise */
  int32_t _pad0;
  double noise_pct[PCG_MAX_NX];    /* 0 = state not perturbed */
  uint64_t seed;
  int64_t env_id0;                 /* global index of local env 0 (keeps noise streams distinct across shards) */

  /* reward */
  int32_t reward;                  /* enum pcg_reward */
  int32_t _pad1;
  double batch_w[PCG_MAX_NX];      /* batch reward: signed r_scale per reward state, 0 elsewhere */
  double penalty;                  /* 1000 in the reference */

  /* affine constraints g = Gx*state + Gu*uk + h  (g > 0 = violated), pcgym.py:560-615.
   * The double un-normalisation quirk of constraint_check is folded into G and h by the host. */
  int32_t n_con;
  int32_t _pad2;
  double con_gx[PCG_MAX_NCON][PCG_MAX_OBS];
  double con_gu[PCG_MAX_NCON][PCG_MAX_NU];
  double con_h[PCG_MAX_NCON];
  uint64_t con_mask_x[PCG_MAX_NCON];  /* nonzero columns of con_gx */
  uint32_t con_mask_u[PCG_MAX_NCON];  /* nonzero columns of con_gu */

  /* PCG_REWARD_OCP */
  double ocp_R;                    /* control-cost weight R */
  double ocp_inv_span[PCG_MAX_NSP];/* 1/(o_hi - o_lo) of each setpoint's state */
  int32_t ocp_mode;                /* bits 0-1: 0 R*u_norm^2, 1 R*(u_norm - u_prev_norm)^2, 2 as 1 but io.u_prev not yet valid;
                                      | PCG_OCP_MOMENTS: the two tracked quantities are derived from the noisy moments */
  int32_t _pad3;
} pcg_desc;

/* Device buffers.  n = n_env, K = k_steps of this call.  Any output pointer may be NULL (not written). */
typedef struct pcg_io {
  double* x;               /* in/out [nx][n]   model state; holds x(t0+K) on completion        */
  const double* a;         /* in     [K][nu_a][n] actions (normalised or physical per flags)   */
  double* a_save;          /* in/out [nu_a][n] a_delta accumulator, or NULL                    */
  const double* sp;        /* in     setpoints; element (t, j, env) at sp[t*sp_st + j*sp_sj + env*sp_se],
                                     t = absolute time index 0..N-1                            */
  int64_t sp_st, sp_sj, sp_se;
  const double* dist;      /* in     disturbances, same addressing with rows = supplied disturbances */
  int64_t d_st, d_sj, d_se;
  const double* z;         /* in     [K][nx][n] standard normals (PCG_NOISE_EXTERNAL), or NULL */
  const double* tail;      /* in     [n_tail][n] observation-only constants, or NULL           */
  double* x_traj;          /* out    [K][nx][n] noise-free state after each step               */
  double* obs;             /* out    [K][n_obs][n] observation (noise, normalisation, mask)    */
  double* reward;          /* out    [K][n]                                                    */
  uint8_t* done;           /* out    [K][n]                                                    */
  uint8_t* violated;       /* out    [K][n]                                                    */
  double* g;               /* out    [K][n_con][n] constraint values, cons_info[:, t0+k+1, :]  */
  double* g0;              /* out    [n_con][n] pre-check at t == 0, cons_info[:, 0, :]        */
  uint8_t* done_state;     /* in/out [n] sticky done flag (self.done), or NULL                 */
  uint8_t* status;         /* in/out [n] OR-accumulated PCG_ST_* bits, or NULL                 */
  double* h_carry;         /* in/out [n] adaptive step size carried between intervals, or NULL */
  int32_t* n_rhs;          /* in/out [n] accumulated RHS evaluations, or NULL                  */
  double* u_prev;          /* in/out [nu_a][n] previous physical action (PCG_REWARD_OCP), or NULL */
  const double* theta;     /* in     [nparam][n] per-environment model parameters (pcg_model_param_name order), or NULL
                                     = desc.theta for every env.  Extension over the reference, where sampled
                                     parameters never reach the dynamics (pcgym.py:301-316 vs integrator.py:28).
                                     Disturbance channels that are not supplied fall back to the env's own theta. */
} pcg_io;

/* ABI version of the loaded library. */
int pcg_abi_version(void);

/* sizeof(pcg_desc) / sizeof(pcg_io) in the library (binding self-check). */
int pcg_sizeof_desc(void);
int pcg_sizeof_io(void);

/* Dimensions of a registry model: states, manipulated inputs, disturbance channels, parameters.
 * Replaces the reference's model.info() lookups (model_classes.py:11-20). */
int pcg_model_dims(int model, int* nx, int* nu, int* nd, int* nparam);
/* 1 if the model provides the analytic Jacobian PCG_SCHEME_RODAS4 needs (cstr, multistage_extraction), else 0 */
int pcg_model_stiff_capable(int model);

/* Name and default of parameter `i` of `model` (reference dataclass field order); NULL if out of range. */
const char* pcg_model_param_name(int model, int i);
int pcg_model_param_default(int model, int i, double* value);

/* Advance every environment by ONE control interval at time index t (make_env.step, pcgym.py:350-500).
 * Equivalent to pcg_rollout(..., t, 1, ...). */
int pcg_step(const pcg_desc* d, const pcg_io* io, int64_t n_env, int64_t t, void* cuda_stream);

/* Advance every environment by k_steps control intervals starting at time index t0 in one launch;
 * the state stays in registers between intervals. */
int pcg_rollout(const pcg_desc* d, const pcg_io* io, int64_t n_env, int64_t t0, int64_t k_steps,
                void* cuda_stream);

/* The measurement-noise transform in isolation (pcgym.py:453-466 draws np.random.normal; here every 32-bit Philox
 * word w gives one draw z = Phi^-1((w + 1/2) 2^-32), see csrc/philox.cuh): z0[i] from w0[i] and z1[i] from w1[i]
 * for i < n, device pointers.  f32 != 0 selects the PCG_NOISE_PHILOX_F32 evaluation (fp32 Box-Muller of the word
 * pair).  Lets a caller check the documented draw exactly. */
int pcg_normal_pairs(const uint32_t* w0, const uint32_t* w1, double* z0, double* z1, int64_t n, int f32,
                     void* cuda_stream);

/* The branch-free fp64 helpers the model right-hand sides use instead of the CUDA library's exp / log / sqrt / divide
 * (csrc/fastmath.cuh), in isolation: y[i] = f(x[i]) for i < n, device pointers.  Lets a caller pin their accuracy
 * (tests: <= 2 ulp against the correctly rounded value over the ranges the models use). */
enum pcg_math_op {
  PCG_MATH_EXP = 0,         /* exp_sat: Estrin degree-10, saturating at 2^+-1000            */
  PCG_MATH_EXP_HORNER = 1,  /* exp_sat_horner                                               */
  PCG_MATH_EXP_TAB = 2,     /* exp_tab_split: 2^(j/256) table + degree-4 (integrator warps) */
  PCG_MATH_LOG = 3,         /* log_tab: log|x|, 128-entry table + degree-7 log1p            */
  PCG_MATH_SQRT = 4,        /* sqrt_fast: MUFU.RSQ64H + Goldschmidt + residual correction   */
  PCG_MATH_RCP = 5,         /* rcp_nr: MUFU.RCP64H + two Newton steps                       */
  PCG_MATH_RCP_O3 = 6,      /* rcp_o3: one third-order correction                           */
  PCG_MATH_COUNT = 7
};
int pcg_math_probe(int op, const double* x, double* y, int64_t n, void* cuda_stream);

/* Number of kernel launches issued by this library since load (diagnostics for bench.py). */
int64_t pcg_launch_count(void);

/* Text for a pcg_status; for PCG_E_CUDA, pcg_last_cuda_error() returns the runtime's message
 * of the most recent failure on the calling thread. */
const char* pcg_strerror(int status);
const char* pcg_last_cuda_error(void);

/* Ask for a kernel family for the calling thread's following launches (enum pcg_kernel); returns the previous
 * request, or PCG_E_RANGE.  A request is honoured only where that family is eligible for the configuration (the usual
 * fall-back order applies otherwise) -- pcg_last_kernel() tells.  Lets the parity tests drive every instantiation
 * at every batch size; production code leaves it at PCG_KERNEL_AUTO. */
int pcg_set_kernel(int kernel);

/* Kernel family the calling thread's most recent pcg_step / pcg_rollout launched (diagnostics: the library
 * picks between the generic, the software-pipelined and the warp-specialised kernel per configuration). */
const char* pcg_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* PCGYM_B200_H */
