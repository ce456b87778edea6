// Device right-hand sides of the registry models (reference src/pcgym/model_classes.py).
//
// Every model is a struct with compile-time dimensions
//   NX   states                       NU  manipulated inputs
//   ND   disturbance input channels   NP  parameters theta (reference dataclass field order)
// a derived-constant block `P` built once per thread from theta, and
//   rhs(P, x[NX], uk[NU+ND], dx[NX])
// where uk = [manipulated inputs | disturbance channels].  The reference switches on
// `u.size` to decide whether the disturbance channels override parameters
// (model_classes.py:48-51, 382-385); here the caller always fills uk[NU..NU+ND) -- with the
// supplied disturbance or with the parameter default -- so the RHS has no branch.
//
// Arithmetic is fp64 throughout.  Expressions are regrouped for fewer DFMA/MUFU
// operations (constants folded, reciprocals hoisted); parity is defined against the
// converged solution of the reference RHS to rtol 1e-6, and the regrouping perturbs
// results at the 1e-15 level.
#pragma once
#include <cmath>
#include <cstdint>

#include "fastmath.cuh"

namespace pcg {

// ---------------------------------------------------------------------------------------
// cstr -- model_classes.py:23-62
// theta: q V rho C deltaHr EA_over_R k0 UA Ti Caf       disturbances: Ti, Caf
#ifndef PCG_CSTR_WSMODEL
#define PCG_CSTR_WSMODEL CstrT
#endif
#ifndef PCG_CSTR_WSMMODEL
#define PCG_CSTR_WSMMODEL CstrT
#endif
struct Cstr {
  static constexpr int NX = 2, NU = 1, ND = 2, NP = 10;
  struct P { double thru, ln_k0, neg_ear, c_heat, c_cool, c_T; };
  PCG_HD static P make(const double* th) {
    P p;
    p.thru = th[0] / th[1];                       // q / V
    p.ln_k0 = log(th[6]);                         // k0 exp(-E/T) = exp(ln k0 - E/T)
    p.neg_ear = -th[5];                           // -EA/R
    p.c_heat = (-th[4]) / (th[2] * th[3]);        // (-deltaHr) / (rho C)
    p.c_cool = th[7] / (th[2] * th[3] * th[1]);   // UA / (rho C V)
    p.c_T = -(p.thru + p.c_cool);
    return p;
  }
  // per-interval constants: the input is held over the control interval, so everything that depends
  // only on uk is hoisted out of the six RHS evaluations
  struct Aux { double feed_c, heat0; };
  PCG_DEV static void prep(const P& p, const double* u, Aux& a) {
    a.feed_c = p.thru * u[2];                             // q/V Caf
    a.heat0 = fma(p.thru, u[1], p.c_cool * u[0]);         // q/V Ti + UA/(rho C V) Tc
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {8, 9}; return t[i]; }      // theta index of each disturbance default
  static constexpr int EPT = 2;                   // environments per thread (ILP 2 on the FP64 pipe)
  static constexpr int MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 7;   // 7 x 64 threads x 144 regs: 65 536 envs in one wave
  static constexpr bool FAST = true;              // has straight-line kernel instantiations (rollout_fast / rollout_ws)
  static constexpr int WS = 1;                    // warp-specialised kernel: integrator/finisher pairs per CTA
  using WsModel = struct PCG_CSTR_WSMODEL;        // RHS variant its integrator warps run (CstrT; -DPCG_CSTR_WSMODEL=CstrH: A/B)
  using WsmModel = struct PCG_CSTR_WSMMODEL;      // the same for the per-SM kernel (CstrT; -DPCG_CSTR_WSMMODEL=CstrW -DPCG_WSM_RING=8 -DPCG_WSM_WARPS=24: A/B)
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double*, double* dx) {
    const double ca = x[0], T = x[1];
    const double rate = exp_sat(fma(p.neg_ear, rcp_nr(T), p.ln_k0)) * ca;        // :52
    dx[0] = fma(-p.thru, ca, a.feed_c) - rate;                                   // :53
    dx[1] = fma(p.c_heat, rate, fma(p.c_T, T, a.heat0));                         // :54-58
  }
  // Rosenbrock interface (scheme "rodas4"): W = s I - J(x) with the analytic Jacobian of the two lines above, kept as
  // its 2 x 2 inverse.  dk/dT = k E/(R T^2).
  static constexpr bool STIFF = true;
  struct Lin { double i00, i01, i10, i11; };
  PCG_DEV static void factor(const P& p, const Aux&, const double* x, const double*, double s, Lin& w) {
    const double ca = x[0], T = x[1];
    const double r = rcp_nr(T);
    const double k = exp_sat(fma(p.neg_ear, r, p.ln_k0));
    const double dr = -p.neg_ear * r * r * k * ca;               // d(k ca)/dT
    const double w00 = s + p.thru + k, w01 = dr;                 // -J00, -J01
    const double w10 = -p.c_heat * k, w11 = s - fma(p.c_heat, dr, p.c_T);
    const double idet = rcp_nr(fma(w00, w11, -(w01 * w10)));
    w.i00 = w11 * idet;  w.i01 = -w01 * idet;
    w.i10 = -w10 * idet; w.i11 = w00 * idet;
  }
  PCG_DEV static void solve(const Lin& w, double* b) {
    const double b0 = b[0], b1 = b[1];
    b[0] = fma(w.i00, b0, w.i01 * b1);
    b[1] = fma(w.i10, b0, w.i11 * b1);
  }
};

// cstr with the Horner-form exponential (integrator warps of the warp-specialised kernel)
struct CstrH : Cstr {
  using WsModel = CstrH;
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double*, double* dx) {
    const double ca = x[0], T = x[1];
    // a NaN T or ca still reaches dx through the linear terms below, so the status check cannot miss it
    const double rate = exp_sat_horner_iscale(fma(p.neg_ear, rcp_nr(T), p.ln_k0)) * ca;
    dx[0] = fma(-p.thru, ca, a.feed_c) - rate;
    dx[1] = fma(p.c_heat, rate, fma(p.c_T, T, a.heat0));
  }
};

// cstr as the integrator warps of the warp-specialised kernels run it: table-driven exponential (exp_tab_split, 8 FP64
// instructions, dependency depth 7) and the third-order reciprocal (3 DFMA): 17 FP64 instructions + 1 MUFU per RHS
// against 25 for CstrH, and a dependency chain about half as long.  The Arrhenius factor comes back as scale * poly;
// scale * ca and c_heat * scale * ca are formed while the polynomial is in flight, so the derivative is one DFMA behind it.
struct CstrT : Cstr {
  using WsModel = CstrT;
  static constexpr bool EXP_TAB = true;           // Aux carries the shared-memory 2^(j/256) table (kExp2Tab)
  struct Aux { double feed_c, heat0; const double* tab; };
  PCG_DEV static void prep(const P& p, const double* u, Aux& a) {
    a.feed_c = p.thru * u[2];
    a.heat0 = fma(p.thru, u[1], p.c_cool * u[0]);
  }
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double*, double* dx) {
    const double ca = x[0], T = x[1];
    double scale, poly;
    exp_tab_split(fma(p.neg_ear, rcp_o3(T), p.ln_k0), a.tab, scale, poly);       // :52  k0 exp(-E/RT) = scale * poly
    const double sc = scale * ca, hs = p.c_heat * sc;
    dx[0] = fma(-sc, poly, fma(-p.thru, ca, a.feed_c));                           // :53
    dx[1] = fma(hs, poly, fma(p.c_T, T, a.heat0));                                // :54-58
  }
};

// A/B variant for the per-SM kernel (rollout_wsm.cuh; build with -DPCG_CSTR_WSMMODEL=CstrW -DPCG_WSM_RING=8
// -DPCG_WSM_WARPS=24 -DPCG_WSM_FIN_REGS=96): the Arrhenius factor k0 exp(-E/RT) comes from a shared-memory table indexed
// DIRECTLY by n = rint(256 y log2 e), y = ln k0 - (E/R)/T, built per CTA for every n a reactor temperature in
// [kTabLo, kTabHi] can produce (exp_wide_split: the integer side of the exponential is one clamp and one address;
// 243 instead of 270 instructions per lane-interval).  The table has EXP_WIDE entries at most (75 KB: the rings shrink
// to 8 slots to make room), which covers 240..800 K for the reference parameters; wide_ok() tells the launcher whether a
// parameter set fits.  A lane whose temperature leaves [kChkLo, kChkHi] at the end of an interval is flagged
// PCG_ST_RANGE.  MEASURED (profiles/README.md r2, same-box A/B): 0.157-0.160 ms against 0.157-0.164 ms for CstrT at
// 65 536 envs -- no gain: at 443 envs per SM the kernel is bound by how many independent dependency chains a scheduler
// holds, not by issue slots -- and +4 % for the adaptive scheme.  Not the default: the table narrows the valid
// temperature range for nothing at the headline.
struct CstrW : Cstr {
  using WsModel = CstrW;
  static constexpr bool EXP_TAB = true;
  static constexpr int EXP_WIDE = 9600;
  static constexpr double kTabLo = 240.0, kTabHi = 800.0, kChkLo = 250.0, kChkHi = 760.0;
  struct P : Cstr::P { double wshift; int n_min, wlast, chk_hi0, chk_span; };
  PCG_HD static P make(const double* th) {
    P p;
    static_cast<Cstr::P&>(p) = Cstr::make(th);
    const double l256 = 369.3299304675746;
    const double n_lo = floor(fma(p.neg_ear, 1.0 / kTabLo, p.ln_k0) * l256) - 2.0;
    const double n_hi = ceil(fma(p.neg_ear, 1.0 / kTabHi, p.ln_k0) * l256) + 2.0;
    const bool sane = n_lo > -1e6 && n_hi < 1e6 && n_hi > n_lo;      // also false for NaN parameters
    p.n_min = sane ? static_cast<int>(n_lo) : 0;
    p.wlast = sane ? static_cast<int>(n_hi - n_lo) : -1;
    p.wshift = 6755399441055744.0 - static_cast<double>(p.n_min);
    p.chk_hi0 = hi_word(kChkLo);
    p.chk_span = hi_word(kChkHi) - p.chk_hi0;
    return p;
  }
  PCG_HD static bool wide_ok(const P& p) { return p.wlast >= 1 && p.wlast < EXP_WIDE && p.neg_ear < 0.0; }
  struct Aux { double feed_c, heat0; const double* tab; };
  PCG_DEV static void prep(const P& p, const double* u, Aux& a) {
    a.feed_c = p.thru * u[2];
    a.heat0 = fma(p.thru, u[1], p.c_cool * u[0]);
  }
  PCG_DEV static void fill_tab(double* tab, const P& p) { exp_wide_fill(tab, p.n_min, p.wlast); }
  // PCG_ST_RANGE when the temperature (or a NaN: high word 0x7ff8....) is outside the checked band; integer compare
  PCG_DEV static unsigned range_flag(const P& p, const double* x) {
    return static_cast<unsigned>(__double2hiint(x[1]) - p.chk_hi0) > static_cast<unsigned>(p.chk_span) ? 4u : 0u;
  }
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double*, double* dx) {
    const double ca = x[0], T = x[1];
    double scale, poly;
    exp_wide_split(fma(p.neg_ear, rcp_o3(T), p.ln_k0), a.tab, p.wshift, static_cast<unsigned>(p.wlast), scale, poly);   // :52
    const double sc = scale * ca, hs = p.c_heat * sc;
    dx[0] = fma(-sc, poly, fma(-p.thru, ca, a.feed_c));                           // :53
    dx[1] = fma(hs, poly, fma(p.c_T, T, a.heat0));                                // :54-58
  }
};

// ---------------------------------------------------------------------------------------
// first_order_system -- model_classes.py:296-343      theta: K tau
// models with the Rosenbrock interface (STIFF, Lin, factor, solve)
template <class M, class = void>
struct is_stiff_capable { static constexpr bool value = false; };
template <class M>
struct is_stiff_capable<M, decltype(void(M::STIFF))> { static constexpr bool value = true; };

struct FirstOrder {
  static constexpr int NX = 1, NU = 1, ND = 0, NP = 2;
  struct P { double K, inv_tau; };
  PCG_HD static P make(const double* th) { return P{th[0], 1.0 / th[1]}; }
  struct Aux {};
  PCG_DEV static void prep(const P&, const double*, Aux&) {}
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  // WS = 0: measured with the warp-specialised kernels the trivial RHS leaves the integrator warps idle and the
  // finishers carry everything: 0.142 ms against 0.118 ms (65 536 envs), 41.8 % against 60.8 % of HBM (1 M envs)
  static constexpr int EPT = 2, MIN_BLOCKS = 8, MIN_BLOCKS_EPT1 = 8, WS = 0;
  static constexpr bool FAST = true;
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    dx[0] = fma(p.K, u[0], -x[0]) * p.inv_tau;                   // :322
  }
};

// ---------------------------------------------------------------------------------------
// multistage_extraction -- model_classes.py:346-430
// theta: Vl Vg m Kla eq_exponent X0 Y6                 disturbances: X0, Y6
// states interleaved X1 Y1 ... X5 Y5
// ExtractionT<5> is the reference model; ExtractionT<10> ("multistage_extraction_10stage") is an EXTENSION with the same
// stage balances (BASELINE.json config 4 words the workload "10-stage, ~20-state"; the reference has no such model).
template <int NS>
struct ExtractionT {
  static constexpr int NSTAGE = NS;
  static constexpr int NX = 2 * NSTAGE, NU = 2, ND = 2, NP = 7;
  struct P { double inv_vl, inv_vg, inv_m, kla_vl, expo; bool square; };
  PCG_HD static P make(const double* th) {
    P p;
    p.inv_vl = 1.0 / th[0];
    p.inv_vg = 1.0 / th[1];
    p.inv_m = 1.0 / th[2];
    p.kla_vl = th[3] * th[0];
    p.expo = th[4];
    p.square = (th[4] == 2.0);
    return p;
  }
  struct Aux {};
  PCG_DEV static void prep(const P&, const double*, Aux&) {}
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {5, 6}; return t[i]; }
  static constexpr int EPT = 1, MIN_BLOCKS = NS <= 5 ? 4 : 2, MIN_BLOCKS_EPT1 = NS <= 5 ? 4 : 2, WS = 0;
  static constexpr bool FAST = NS <= 5;           // the straight-line kernel keeps two intervals of state in registers
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double L = u[0], G = u[1], X0 = u[2], Y6 = u[3];
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) {
      const double X = x[2 * i], Y = x[2 * i + 1];
      const double ye = p.square ? Y * Y : pow(Y, p.expo);       // :387-391
      const double Q = p.kla_vl * fma(-ye, p.inv_m, X);          // :393-397
      const double xin = (i == 0) ? X0 : x[2 * i - 2];
      const double yin = (i == NSTAGE - 1) ? Y6 : x[2 * i + 3];
      dx[2 * i] = p.inv_vl * fma(L, xin - X, -Q);                // :399-410
      dx[2 * i + 1] = p.inv_vg * fma(G, yin - Y, Q);
    }
  }
  // Lane-split form (rollout_lanes.cuh): lane s of an env's group owns stage s, xl = (X_s, Y_s).  The liquid comes from
  // the stage on the left, the gas from the stage on the right: one 64-bit shuffle each, executed by the whole warp.  The
  // two ends of the column read the FEEDS (X0, Y6) from a spare lane of the warp that holds them as its "state" with
  // all-zero coefficients (its derivative is exactly 0, so it stays put through every stage): no per-evaluation select.
  // LaneC = the interval's coefficients (inputs folded in), optionally pre-multiplied by the step so that the evaluation
  // returns h f directly.  Quadratic equilibrium only (the reference's default eq_exponent = 2: lanes_ready()).
  static constexpr int LANES = NS;
  PCG_HD static bool lanes_ready(const P& p) { return p.square; }
  struct LaneC { double cl, cg, ql, qg; };
  PCG_DEV static void lane_prep(const P& p, const double* u, bool live, LaneC& c) {
    c.cl = live ? p.inv_vl * u[0] : 0.0;                         // L / Vl
    c.cg = live ? p.inv_vg * u[1] : 0.0;                         // G / Vg
    c.ql = live ? p.inv_vl * p.kla_vl : 0.0;                     // Kla
    c.qg = live ? p.inv_vg * p.kla_vl : 0.0;                     // Kla Vl / Vg
  }
  PCG_DEV static void lane_feed(const double* u, double (&xl)[2]) { xl[0] = u[2]; xl[1] = u[3]; }   // (X0, Y6)
  PCG_DEV static void rhs_lane(const P& p, const LaneC& c, const double (&xl)[2], int src_l, int src_r, double (&dxl)[2]) {
    const double X = xl[0], Y = xl[1];
    const double xin = __shfl_sync(0xffffffffu, X, src_l), yin = __shfl_sync(0xffffffffu, Y, src_r);
    const double q = fma(-(Y * Y), p.inv_m, X);                  // X - Y^2 / m   (:387-397)
    dxl[0] = fma(-c.ql, q, c.cl * (xin - X));                    // :399-410
    dxl[1] = fma(c.qg, q, c.cg * (yin - Y));
  }
  // Rosenbrock interface (scheme "rodas4"): with the states interleaved X1 Y1 X2 Y2 ... the Jacobian of the column is
  // pentadiagonal (X_i sees X_{i-1} two columns to the left and Y_i one to the right; Y_i sees X_i one to the left and
  // Y_{i+1} two to the right).  W = s I - J is a column-diagonally-dominant Z-matrix (the stage balances conserve mass),
  // so the banded LU needs no pivoting; it is kept as unit-lower multipliers, the upper band and reciprocal pivots.
  static constexpr bool STIFF = true;
  struct Lin { double l1[NX], l2[NX], u1[NX], u2[NX], ip[NX]; };
  PCG_DEV static void factor(const P& p, const Aux&, const double* x, const double* u, double s, Lin& w) {
    const double L = u[0], G = u[1];
    double d[NX], a1[NX], a2[NX], b1[NX], b2[NX];        // diagonal, sub-1, sub-2, super-1, super-2 of W
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) {
      const double Y = x[2 * i + 1];
      const double dye = (p.square ? 2.0 * Y : p.expo * pow(Y, p.expo - 1.0)) * p.inv_m;     // d(Y^e / m)/dY
      const double q = p.kla_vl * dye;
      d[2 * i] = s + p.inv_vl * (L + p.kla_vl);
      a1[2 * i] = 0.0;  a2[2 * i] = -(p.inv_vl * L);     // X_{i-1} (absent for i == 0: never read)
      b1[2 * i] = -(p.inv_vl * q);  b2[2 * i] = 0.0;
      d[2 * i + 1] = s + p.inv_vg * (G + q);
      a1[2 * i + 1] = -(p.inv_vg * p.kla_vl);  a2[2 * i + 1] = 0.0;
      b1[2 * i + 1] = 0.0;  b2[2 * i + 1] = -(p.inv_vg * G);   // Y_{i+1} (absent for the last stage: never read)
    }
#pragma unroll
    for (int k = 0; k < NX; ++k) {
      w.ip[k] = rcp_nr(d[k]);
      w.u1[k] = b1[k];
      w.u2[k] = b2[k];
      if (k + 1 < NX) {
        const double m = a1[k + 1] * w.ip[k];
        w.l1[k + 1] = m;
        d[k + 1] = fma(-m, b1[k], d[k + 1]);
        b1[k + 1] = fma(-m, b2[k], b1[k + 1]);
      }
      if (k + 2 < NX) {
        const double m = a2[k + 2] * w.ip[k];
        w.l2[k + 2] = m;
        a1[k + 2] = fma(-m, b1[k], a1[k + 2]);
        d[k + 2] = fma(-m, b2[k], d[k + 2]);
      }
    }
  }
  PCG_DEV static void solve(const Lin& w, double* b) {
#pragma unroll
    for (int i = 1; i < NX; ++i) {
      double v = fma(-w.l1[i], b[i - 1], b[i]);
      if (i >= 2) v = fma(-w.l2[i], b[i - 2], v);
      b[i] = v;
    }
#pragma unroll
    for (int i = NX - 1; i >= 0; --i) {
      double v = b[i];
      if (i + 1 < NX) v = fma(-w.u1[i], b[i + 1], v);
      if (i + 2 < NX) v = fma(-w.u2[i], b[i + 2], v);
      b[i] = v * w.ip[i];
    }
  }
};
using Extraction = ExtractionT<5>;
using Extraction10 = ExtractionT<10>;

// ---------------------------------------------------------------------------------------
// four_tank -- model_classes.py:864-931
// theta: g gamma_1 gamma_2 k1 k2 a1 a2 a3 a4 A1 A2 A3 A4   (disturbances ["None"]: inert)
struct FourTank {
  static constexpr int NX = 4, NU = 2, ND = 0, NP = 13;
  struct P { double g2, o11, o13, o22, o24, o33, o44, p11, p22, p32, p41; };
  PCG_HD static P make(const double* th) {
    P p;
    p.g2 = 2.0 * th[0];
    const double A1 = th[9], A2 = th[10], A3 = th[11], A4 = th[12];
    p.o11 = -th[5] / A1;  p.o13 = th[7] / A1;
    p.o22 = -th[6] / A2;  p.o24 = th[8] / A2;
    p.o33 = -th[7] / A3;  p.o44 = -th[8] / A4;
    p.p11 = (th[1] * th[3]) / A1;
    p.p22 = (th[2] * th[4]) / A2;
    p.p32 = ((1.0 - th[2]) * th[4]) / A3;
    p.p41 = ((1.0 - th[1]) * th[3]) / A4;
    return p;
  }
  struct Aux {};
  PCG_DEV static void prep(const P&, const double*, Aux&) {}
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  static constexpr int EPT = 1, MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 4, WS = 0;
  static constexpr bool FAST = true;
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    // sqrt_fast: branch-free (the four roots interleave in one basic block), NaN for h < 0 as the reference's np.sqrt
    const double s1 = sqrt_fast(p.g2 * x[0]), s2 = sqrt_fast(p.g2 * x[1]);
    const double s3 = sqrt_fast(p.g2 * x[2]), s4 = sqrt_fast(p.g2 * x[3]);
    dx[0] = fma(p.o11, s1, fma(p.o13, s3, p.p11 * u[0]));          // :906
    dx[1] = fma(p.o22, s2, fma(p.o24, s4, p.p22 * u[1]));          // :907
    dx[2] = fma(p.o33, s3, p.p32 * u[1]);                          // :908
    dx[3] = fma(p.o44, s4, p.p41 * u[0]);                          // :909
  }
};

// ---------------------------------------------------------------------------------------
// crystallization -- model_classes.py:1232-1345
// theta: ka kb kc kd kg k1 k2 a b alfa ro      disturbances ka, kg, UA are declared (:1342)
// but the RHS reads only u[0] (:1299), so they are carried in uk and ignored.
struct Crystallization {
  static constexpr int NX = 7, NU = 1, ND = 3, NP = 11;
  struct P { double ka, kb, kc, kd, kg, k1, k2, a, b, half_ro_alfa; };
  PCG_HD static P make(const double* th) {
    return P{th[0], th[1], th[2], th[3], th[4], th[5], th[6], th[7], th[8], 0.5 * th[10] * th[9]};
  }
  // per-interval constants: the jacket temperature is held over the control interval, so the solubility polynomial and
  // the two Arrhenius exponents' temperature terms are hoisted out of the right-hand side
  struct Aux { double ceq, b_t, g_t; };
  PCG_DEV static void prep(const P& p, const double* u, Aux& a) {
    const double TK = u[0] + 273.15;
    const double inv_TK = 1.0 / TK;
    a.ceq = fma(TK, fma(-0.00292874, TK, 3.579165), -686.2686);                 // :1302
    a.b_t = p.kb * inv_TK;
    a.g_t = p.k1 * inv_TK;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0, 4, -1}; return t[i]; }   // "UA" is not a parameter of this model
  static constexpr int EPT = 1, MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 4, WS = 0;
  static constexpr bool FAST = true;
  // Branch-free arithmetic (fastmath.cuh: table log, exp_sat, Newton reciprocal, Goldschmidt sqrt; each within ~1 ulp):
  // the library versions carry slow-path branches and ~330 instructions per evaluation, this form ~150.
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double*, double* dx) {
    const double mu0 = x[0], mu1 = x[1], mu2 = x[2], mu3 = x[3], conc = x[4];
    const double S = fma(conc, 1e3, -a.ceq);                                    // :1303
    const double lnS = log_tab(S);              // (S^2)^(k/2) = |S|^k = exp(k ln|S|); log_tab takes |x|
    const double lnM = log_tab(mu3);
    const double B0 = p.ka * exp_sat(fma(p.kc, lnS, fma(p.kd, lnM, a.b_t)));    // :1304
    const double G = p.kg * exp_sat(fma(p.k2, lnS, a.g_t));                     // :1305
    const double d0 = B0;
    const double d1 = G * fma(p.a, mu0, p.b * mu1 * 1e-4) * 1e4;                   // :1308
    const double d2 = 2.0 * G * fma(p.a * mu1, 1e-4, p.b * mu2 * 1e-8) * 1e8;      // :1309
    const double m23 = fma(p.a * mu2, 1e-8, p.b * mu3 * 1e-12);
    const double d3 = 3.0 * G * m23 * 1e12;                                        // :1310
    const double dc = -p.half_ro_alfa * G * m23;                                   // :1311
    const double mu1sq = mu1 * mu1;
    const double CV = sqrt_fast(mu2 * mu0 * rcp_nr(mu1sq) - 1.0);                  // :1313
    const double num = fma(fma(d2, mu0, mu2 * d0), mu1sq, -(mu2 * mu0 * 2.0 * mu1 * d1));
    const double dCV = rcp_nr(fma(2.0, CV, 1e-10)) * num * rcp_nr(fma(mu1sq, mu1sq, 1e-10));   // :1314
    const double dLn = fma(d1, mu0, -(mu1 * d0)) * rcp_nr(fma(mu0, mu0, 1e-10));   // :1315
    dx[0] = d0; dx[1] = d1; dx[2] = d2; dx[3] = d3; dx[4] = dc; dx[5] = dCV; dx[6] = dLn;
  }
};

// ---------------------------------------------------------------------------------------
// biofilm_reactor -- model_classes.py:1046-1155
// theta: V Va Kla m eq_exponent O_air vm_1 vm_2 K1 K2 KO_1 KO_2    inputs F Fr S1_F S2_F S3_F
struct Biofilm {
  static constexpr int NX = 16, NU = 5, ND = 0, NP = 12;
  struct P { double inv_v, inv_va, kla, o_eq, vm1, vm2, K1, K2, KO1, KO2; };
  PCG_HD static P make(const double* th) {
    P p;
    p.inv_v = 1.0 / th[0];
    p.inv_va = 1.0 / th[1];
    p.kla = th[2];
    p.o_eq = ((th[4] == 1.0) ? th[5] : pow(th[5], th[4])) / th[3];   // O_air^e / m  (:1112)
    p.vm1 = th[6]; p.vm2 = th[7]; p.K1 = th[8]; p.K2 = th[9]; p.KO1 = th[10]; p.KO2 = th[11];
    return p;
  }
  struct Aux {};
  PCG_DEV static void prep(const P&, const double*, Aux&) {}
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  static constexpr int EPT = 1, MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 4, WS = 0;
  static constexpr bool FAST = true;
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double F = u[0], Fr = u[1];
    const double frv = Fr * p.inv_v, fra = Fr * p.inv_va, fa = F * p.inv_va;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double S1 = x[4 * k], S2 = x[4 * k + 1], S3 = x[4 * k + 2], O = x[4 * k + 3];
      const int up = (k == 0) ? 12 : 4 * (k - 1);                  // absorber feeds stage 1
      // Monod terms (:1091-1100); one divide per reaction
      const double r1 = (p.vm1 * S1 * O) / ((p.K1 + S1) * (p.KO1 + O));
      const double r2 = (p.vm2 * S2 * O) / ((p.K2 + S2) * (p.KO2 + O));
      dx[4 * k] = fma(frv, x[up] - S1, r1);                        // - rs1, rs1 = -r1
      dx[4 * k + 1] = fma(frv, x[up + 1] - S2, r2 - r1);           // - rs2, rs2 = r1 - r2
      dx[4 * k + 2] = fma(frv, x[up + 2] - S3, -r2);               // - rs3, rs3 = r2
      dx[4 * k + 3] = fma(frv, x[up + 3] - O, fma(3.5, r1, 1.1 * r2));   // - ro, ro = -3.5 r1 - 1.1 r2
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
      dx[12 + j] = fma(fra, x[8 + j] - x[12 + j], fa * (u[2 + j] - x[12 + j]));   // :1127-1129
    dx[15] = fma(fra, x[11] - x[15], p.kla * (p.o_eq - x[15]));                   // :1130
  }
};

}  // namespace pcg
