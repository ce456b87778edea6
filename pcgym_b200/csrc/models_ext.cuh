// Device right-hand sides of the remaining registry models (reference src/pcgym/pcgym.py:128-148).
// Same contract as models.cuh.  These run in the generic control-interval kernel only (FAST = false): they are the
// breadth of the registry, not the benchmarked path, so the arithmetic follows the reference expression by
// expression (IEEE divides, library sqrt / pow) apart from hoisted parameter combinations.
//
// Models the reference declares WITHOUT inputs (invariant_batch, coupled_oscillator) carry one ignored input row:
// the kernels index actions as a[K][nu][n] with nu >= 1, and the reference would equally need a non-empty
// `a_space` to build its action space.
#pragma once
#include "models.cuh"

namespace pcg {

#define PCG_EXT_TRAITS                                                                      \
  static constexpr int EPT = 1, MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 4, WS = 0;                \
  static constexpr bool FAST = false;                                                       \
  struct Aux {};                                                                            \
  PCG_DEV static void prep(const P&, const double*, Aux&) {}

// ---------------------------------------------------------------------------------------
// complex_cstr -- model_classes.py:65-125     A -> 2B -> C, states Ca Cb Cc T
// theta: q V rho C deltaHr1 EA1_over_R k01 deltaHr2 EA2_over_R k02 UA Ti Caf     disturbances: Ti, Caf
struct ComplexCstr {
  static constexpr int NX = 4, NU = 1, ND = 2, NP = 13;
  struct P { double thru, ln_k01, ea1, ln_k02, ea2, h1, h2, c_cool; };
  PCG_HD static P make(const double* th) {
    P p;
    p.thru = th[0] / th[1];
    p.ln_k01 = log(th[6]);  p.ea1 = -th[5];
    p.ln_k02 = log(th[9]);  p.ea2 = -th[8];
    p.h1 = (-th[4]) / (th[2] * th[3]);
    p.h2 = (-th[7]) / (th[2] * th[3]);
    p.c_cool = th[10] / (th[2] * th[3] * th[1]);
    return p;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {11, 12}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double ca = x[0], cb = x[1], cc = x[2], T = x[3];
    const double rT = rcp_nr(T);
    const double r1 = exp_sat(fma(p.ea1, rT, p.ln_k01)) * ca;                   // :104
    const double r2 = exp_sat(fma(p.ea2, rT, p.ln_k02)) * cb;                   // :105
    dx[0] = p.thru * (u[2] - ca) - r1;                                          // :107
    dx[1] = fma(2.0, r1, -p.thru * cb) - r2;                                    // :108
    dx[2] = fma(-p.thru, cc, r2);                                               // :109
    dx[3] = fma(p.thru, u[1] - T, fma(p.h1, r1, p.h2 * r2)) + p.c_cool * (u[0] - T);   // :111-116
  }
};

// ---------------------------------------------------------------------------------------
// nonsmooth_control -- model_classes.py:509-558      theta: a_11 a_12 a_21 a_22 b_1 b_2
struct Nonsmooth {
  static constexpr int NX = 2, NU = 1, ND = 0, NP = 6;
  struct P { double a11, a12, a21, a22, b1, b2; };
  PCG_HD static P make(const double* th) { return P{th[0], th[1], th[2], th[3], th[4], th[5]}; }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    dx[0] = fma(p.a11, x[0], fma(p.a12, x[1], p.b1 * u[0]));                    // :536
    dx[1] = fma(p.a21, x[0], fma(p.a22, x[1], p.b2 * u[0]));                    // :537
  }
};

// ---------------------------------------------------------------------------------------
// cstr_series_recycle -- model_classes.py:611-679    states C1 T1 C2 T2, inputs F L Tc1 Tc2
// theta: C_O T_O V1 V2 U1A1 U2A2 rho cp k E deltaH R
struct CstrSeries {
  static constexpr int NX = 4, NU = 4, ND = 0, NP = 12;
  struct P { double co_v1, to_v1, inv_v1, inv_v2, ua1, ua2, k, heat, neg_e_r; };
  PCG_HD static P make(const double* th) {
    P p;
    p.co_v1 = th[0] / th[2];
    p.to_v1 = th[1] / th[2];
    p.inv_v1 = 1.0 / th[2];
    p.inv_v2 = 1.0 / th[3];
    p.ua1 = th[4] / (th[2] * th[6] * th[7]);
    p.ua2 = th[5] / (th[3] * th[6] * th[7]);
    p.k = th[8];
    p.heat = (th[8] * (-th[10])) / (th[6] * th[7]);
    p.neg_e_r = -th[9] / th[11];
    return p;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double C1 = x[0], T1 = x[1], C2 = x[2], T2 = x[3];
    const double F = u[0], L = u[1], Tc1 = u[2], Tc2 = u[3], FL = F + L;
    const double e1 = exp_sat(p.neg_e_r * rcp_nr(T1)), e2 = exp_sat(p.neg_e_r * rcp_nr(T2));
    dx[0] = p.co_v1 * F + p.inv_v1 * L * C2 - p.inv_v1 * FL * C1 - p.k * C1 * e1;                                   // :650
    dx[1] = p.to_v1 * F + p.inv_v1 * L * T2 - p.ua1 * (T1 - Tc1) - p.inv_v1 * FL * T1 + p.heat * C1 * e1;           // :651
    dx[2] = p.inv_v2 * FL * (C1 - C2) - p.k * C2 * e2;                                                              // :652
    dx[3] = p.inv_v2 * FL * (T1 - T2) - p.ua2 * (T2 - Tc2) + p.heat * C2 * e2;                                      // :653
  }
};

// ---------------------------------------------------------------------------------------
// distillation_column -- model_classes.py:682-760    states X0 X1 X2 X3 Xf X4 X5 X6 Xb, inputs R F
// theta: D q alpha X_feed M0 Mb M
struct Distillation {
  static constexpr int NX = 9, NU = 2, ND = 0, NP = 7;
  struct P { double D, q, alpha, am1, xfeed, inv_m0, inv_mb, inv_m; };
  PCG_HD static P make(const double* th) {
    return P{th[0], th[1], th[2], th[2] - 1.0, th[3], 1.0 / th[4], 1.0 / th[5], 1.0 / th[6]};
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double R = u[0], F = u[1];
    const double L = R * p.D, V = (R + 1.0) * p.D;                              // :710-711
    const double Ld = fma(p.q, F, L), Vd = fma(1.0 - p.q, F, V), W = F - p.D;   // :712-714
    double y[NX];
#pragma unroll
    for (int i = 1; i < NX; ++i) y[i] = (p.alpha * x[i]) / fma(p.am1, x[i], 1.0);   // :715-722 (Y of X1..Xb)
    dx[0] = p.inv_m0 * (V * y[1] - (L + p.D) * x[0]);                           // :724
    dx[1] = p.inv_m * fma(L, x[0] - x[1], V * (y[2] - y[1]));
    dx[2] = p.inv_m * fma(L, x[1] - x[2], V * (y[3] - y[2]));
    dx[3] = p.inv_m * fma(L, x[2] - x[3], V * (y[4] - y[3]));
    dx[4] = p.inv_m * (L * x[3] - Ld * x[4] + Vd * y[5] - V * y[4] + F * p.xfeed);   // :728
    dx[5] = p.inv_m * fma(Ld, x[4] - x[5], Vd * (y[6] - y[5]));
    dx[6] = p.inv_m * fma(Ld, x[5] - x[6], Vd * (y[7] - y[6]));
    dx[7] = p.inv_m * fma(Ld, x[6] - x[7], Vd * (y[8] - y[7]));
    dx[8] = p.inv_mb * (Ld * x[7] - W * x[8] - Vd * y[8]);                      // :732
  }
};

// ---------------------------------------------------------------------------------------
// multistage_extraction_reactive -- model_classes.py:763-861   5 stages x (XA, YA, YB, YC), inputs L G
// theta: Vl Vg m Kla k eq_exponent XA0 YA6 YB6 YC6
struct ExtractionReactive {
  static constexpr int NSTAGE = 5;
  static constexpr int NX = 4 * NSTAGE, NU = 2, ND = 0, NP = 10;
  struct P { double inv_vl, inv_vg, vg, inv_m, kla_vl, k, expo, xa0, ya6, yb6, yc6; bool square; };
  PCG_HD static P make(const double* th) {
    P p;
    p.inv_vl = 1.0 / th[0];  p.inv_vg = 1.0 / th[1];  p.vg = th[1];
    p.inv_m = 1.0 / th[2];
    p.kla_vl = th[3] * th[0];
    p.k = th[4];
    p.expo = th[5];
    p.square = (th[5] == 2.0);
    p.xa0 = th[6];  p.ya6 = th[7];  p.yb6 = th[8];  p.yc6 = th[9];
    return p;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double L = u[0], G = u[1];
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) {
      const double XA = x[4 * i], YA = x[4 * i + 1], YB = x[4 * i + 2], YC = x[4 * i + 3];
      const double eq = (p.square ? YA * YA : pow(YA, p.expo)) * p.inv_m;       // :797-801
      const double Q = p.kla_vl * (XA - eq);                                    // :803-807
      const double r = p.k * YA * YB;                                           // :809-813
      const bool lastg = (i == NSTAGE - 1);
      const double xin = (i == 0) ? p.xa0 : x[4 * i - 4];
      const double ya_in = lastg ? p.ya6 : x[4 * i + 5];
      const double yb_in = lastg ? p.yb6 : x[4 * i + 6];
      const double yc_in = lastg ? p.yc6 : x[4 * i + 7];
      dx[4 * i] = p.inv_vl * fma(L, xin - XA, -Q);                              // :816-839
      dx[4 * i + 1] = p.inv_vg * (fma(G, ya_in - YA, Q) - r * p.vg);
      dx[4 * i + 2] = p.inv_vg * fma(G, yb_in - YB, -(r * p.vg));
      dx[4 * i + 3] = p.inv_vg * fma(G, yc_in - YC, r * p.vg);
    }
  }
};

// ---------------------------------------------------------------------------------------
// heat_exchanger -- model_classes.py:935-1043    8 cells x (Tt, Tm, Ts), inputs Ft Fs Tt0 Ts9
// theta: Utm Usm L Dt Dm Ds cpt cpm cps rhot rhom rhos
struct HeatExchanger {
  static constexpr int NCELL = 8;
  static constexpr int NX = 3 * NCELL, NU = 4, ND = 0, NP = 12;
  struct P { double ut_at, us_am, ct, cm, cs, cpt, cps; };
  PCG_HD static P make(const double* th) {
    constexpr double pi = 3.141592653589793;
    const double L = th[2], Dt = th[3], Dm = th[4], Ds = th[5];
    const double Vt = L * pi * (Dt * Dt), At = L * pi * Dt;                     // :970-974
    const double Vm = L * pi * (Dm * Dm - Dt * Dt), Am = L * pi * Dm;
    const double Vs = L * pi * (Ds * Ds - Dm * Dm);
    P p;
    p.ut_at = th[0] * At;
    p.us_am = th[1] * Am;
    p.ct = 1.0 / (th[6] * th[9] * Vt);
    p.cm = 1.0 / (th[7] * th[10] * Vm);
    p.cs = 1.0 / (th[8] * th[11] * Vs);
    p.cpt = th[6];
    p.cps = th[8];
    return p;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double Ft = u[0], Fs = u[1], Tt0 = u[2], Ts9 = u[3];
#pragma unroll
    for (int i = 0; i < NCELL; ++i) {
      const double Tt = x[3 * i], Tm = x[3 * i + 1], Ts = x[3 * i + 2];
      const double Qt = p.ut_at * (Tt - Tm), Qm = p.us_am * (Tm - Ts);          // :975-990
      const double tt_in = (i == 0) ? Tt0 : x[3 * i - 3];
      const double ts_in = (i == NCELL - 1) ? Ts9 : x[3 * i + 5];
      dx[3 * i] = p.ct * (Ft * p.cpt * (tt_in - Tt) - Qt);                      // :992-1015
      dx[3 * i + 1] = p.cm * (Qt - Qm);
      dx[3 * i + 2] = p.cs * (Fs * p.cps * (ts_in - Ts) + Qm);
    }
  }
};

// ---------------------------------------------------------------------------------------
// polymerisation_reactor -- model_classes.py:1158-1229   states T M I, inputs F Tf Mf If
// theta: Ap Ad At Ep_over_R Ed_over_R Et_over_R f V deltaHp rho cp
struct Polymerisation {
  static constexpr int NX = 3, NU = 4, ND = 0, NP = 11;
  struct P { double ln_ap, ln_ad, ln_at, ep, ed, et, f, inv_v, heat; };
  PCG_HD static P make(const double* th) {
    return P{log(th[0]), log(th[1]), log(th[2]), -th[3], -th[4], -th[5], th[6], 1.0 / th[7],
             (-th[8]) / (th[9] * th[10])};
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double T = x[0], M = x[1], I = x[2];
    const double fv = u[0] * p.inv_v;
    const double rT = rcp_nr(T);
    const double kp = exp_sat(fma(p.ep, rT, p.ln_ap));                          // :1201-1203
    const double kd = exp_sat(fma(p.ed, rT, p.ln_ad));
    const double kt = exp_sat(fma(p.et, rT, p.ln_at));
    const double ri = 2.0 * p.f * kd * I;                                       // :1204
    const double rp = kp * sqrt((p.f * kd * I) / kt);                           // :1205 (NaN for I < 0, as the reference)
    dx[0] = fma(fv, u[1] - T, p.heat * rp);                                     // :1207-1209
    dx[1] = fma(fv, u[2] - M, -rp);
    dx[2] = fma(fv, u[3] - I, -ri);
  }
};

// ---------------------------------------------------------------------------------------
// photobioreactor (photo_production) -- model_classes.py:433-506   states c_x c_N c_q, inputs I F_N
// theta: u_m u_d Y_NX k_m k_d k_sq K_Nq k_iq k_s k_i k_N
struct PhotoProduction {
  static constexpr int NX = 3, NU = 2, ND = 0, NP = 11;
  struct P { double u_m, u_d, y_nx, k_m, k_d, k_sq, K_Nq, k_iq, k_s, k_i, k_N; };
  PCG_HD static P make(const double* th) {
    return P{th[0], th[1], th[2], th[3], th[4], th[5], th[6], th[7], th[8], th[9], th[10]};
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  static constexpr int EPT = 1, MIN_BLOCKS = 4, MIN_BLOCKS_EPT1 = 4, WS = 0;
  static constexpr bool FAST = false;
  struct Aux { double gx, gq; };                 // light-limitation factors: functions of the held input only
  PCG_DEV static void prep(const P& p, const double* u, Aux& a) {
    const double I = u[0];
    a.gx = p.u_m * I / (I + p.k_s + (I * I / p.k_i));                           // :480
    a.gq = p.k_m * I / (I + p.k_sq + (I * I / p.k_iq));                         // :482
  }
  PCG_DEV static void rhs(const P& p, const Aux& a, const double* x, const double* u, double* dx) {
    const double cx = x[0], cN = x[1], cq = x[2];
    const double growth = a.gx * cx * cN / (cN + p.k_N);
    dx[0] = growth - p.u_d * cx;                                                // :480
    dx[1] = fma(-p.y_nx, growth, u[1]);                                         // :481
    dx[2] = a.gq * cx - (p.k_d * cq) / (cN + p.K_Nq);                           // :482
  }
};

// ---------------------------------------------------------------------------------------
// invariant_batch -- model_classes.py:268-293      states xA xB xC xD, no inputs     theta: k1f k1r k2f k2r
struct InvariantBatch {
  static constexpr int NX = 4, NU = 1, ND = 0, NP = 4;
  struct P { double k1f, k1r, k2f, k2r; };
  PCG_HD static P make(const double* th) { return P{th[0], th[1], th[2], th[3]}; }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double*, double* dx) {
    const double q1 = p.k1f * x[0] * x[1] - p.k1r * x[2];
    const double q2 = p.k2f * x[0] * x[2] - p.k2r * x[3];
    dx[0] = -q1 - q2;                                                           // :285-288
    dx[1] = -q1;
    dx[2] = q1 - q2;
    dx[3] = q2;
  }
};

// ---------------------------------------------------------------------------------------
// batch -- model_classes.py:222-263      states Ca Cb Cc T, input Tc
// theta: k01 k02 EA1 EA2 R dH1 dH2 rho Cp UA V
struct BatchReactor {
  static constexpr int NX = 4, NU = 1, ND = 0, NP = 11;
  struct P { double ln_k01, ln_k02, e1, e2, h1, h2, cool; };
  PCG_HD static P make(const double* th) {
    P p;
    p.ln_k01 = log(th[0]);  p.ln_k02 = log(th[1]);
    p.e1 = -th[2] / th[4];  p.e2 = -th[3] / th[4];
    p.h1 = th[5] / (th[7] * th[8]);  p.h2 = th[6] / (th[7] * th[8]);
    p.cool = th[9] / (th[7] * th[8] * th[10]);
    return p;
  }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double T = x[3];
    const double rT = rcp_nr(T);
    const double r1 = exp_sat(fma(p.e1, rT, p.ln_k01)) * x[0];                  // :249
    const double r2 = exp_sat(fma(p.e2, rT, p.ln_k02)) * x[1];                  // :250
    dx[0] = -r1;
    dx[1] = fma(2.0, r1, -r2);
    dx[2] = r2;
    dx[3] = fma(p.cool, u[0] - T, -fma(p.h1, r1, p.h2 * r2));                   // :254-257
  }
};

// ---------------------------------------------------------------------------------------
// coupled_oscillator -- model_classes.py:186-216   ring of N = 10 masses: positions then momenta, no inputs
// theta: N k m   (the state count is 2 N: only the default N = 10 has a device instantiation)
struct Oscillators {
  static constexpr int NOSC = 10;
  static constexpr int NX = 2 * NOSC, NU = 1, ND = 0, NP = 3;
  struct P { double k, inv_m; };
  PCG_HD static P make(const double* th) { return P{th[1], 1.0 / th[2]}; }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double*, double* dx) {
#pragma unroll
    for (int i = 0; i < NOSC; ++i) {
      const double left = x[(i + NOSC - 1) % NOSC], right = x[(i + 1) % NOSC];
      dx[i] = x[NOSC + i] * p.inv_m;                                            // :208
      dx[NOSC + i] = -p.k * (2.0 * x[i] - left - right);                        // :210-213
    }
  }
};

// ---------------------------------------------------------------------------------------
// disease -- model_classes.py:156-180      states S I R, input u (vaccination rate)     theta: beta gamma
struct Disease {
  static constexpr int NX = 3, NU = 1, ND = 0, NP = 2;
  struct P { double beta, gamma; };
  PCG_HD static P make(const double* th) { return P{th[0], th[1]}; }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double S = x[0], I = x[1], inf = p.beta * S * I, vac = u[0] * S;
    dx[0] = -inf - vac;                                                         // :172-174
    dx[1] = inf - p.gamma * I;
    dx[2] = fma(p.gamma, I, vac);
  }
};

// ---------------------------------------------------------------------------------------
// hydraulic_tank -- model_classes.py:128-150     states q1 q2, input u     theta: D
struct HydraulicTank {
  static constexpr int NX = 2, NU = 1, ND = 0, NP = 1;
  struct P { double D; };
  PCG_HD static P make(const double* th) { return P{th[0]}; }
  PCG_HD static constexpr int dpar(int i) { constexpr int t[] = {0}; return t[i]; }
  PCG_EXT_TRAITS
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
    const double flow = p.D * (x[0] - x[1]);
    dx[0] = u[0] - flow;                                                        // :144-145
    dx[1] = flow - u[0];
  }
};

// ---------------------------------------------------------------------------------------
// Affine custom model  dx = A x + B u + c  (env_params["custom_model"], pcgym.py:150-153; the reference's example
// tests/environment/test_make_env_custom_model.py:7-26 is of this form).  The host probes the user's callable for
// affinity and passes theta = [A row-major | B row-major | c].  Tsit5 instantiations only.
template <int NX_, int NU_>
struct Affine {
  static constexpr int NX = NX_, NU = NU_, ND = 0, NP = NX_ * NX_ + NX_ * NU_ + NX_;
  struct P { double A[NX_][NX_], B[NX_][NU_], c[NX_]; };
  PCG_HD static P make(const double* th) {
    P p;
    for (int i = 0; i < NX; ++i) {
      for (int j = 0; j < NX; ++j) p.A[i][j] = th[i * NX + j];
      for (int j = 0; j < NU; ++j) p.B[i][j] = th[NX * NX + i * NU + j];
      p.c[i] = th[NX * NX + NX * NU + i];
    }
    return p;
  }
  PCG_HD static constexpr int dpar(int) { return -1; }
  PCG_EXT_TRAITS
  static constexpr bool TSIT5_ONLY = true;
  PCG_DEV static void rhs(const P& p, const Aux&, const double* x, const double* u, double* dx) {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double acc = p.c[i];
#pragma unroll
      for (int j = 0; j < NX; ++j) acc = fma(p.A[i][j], x[j], acc);
#pragma unroll
      for (int j = 0; j < NU; ++j) acc = fma(p.B[i][j], u[j], acc);
      dx[i] = acc;
    }
  }
};

#undef PCG_EXT_TRAITS

}  // namespace pcg
