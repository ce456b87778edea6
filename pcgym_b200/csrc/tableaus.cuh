// Butcher tableaus for the explicit Runge-Kutta schemes (autonomous RHS, so no c vector).
//
//   S      stages that define the solution
//   a(i,j) stage coefficients, i in [1, S]  (row S is the FSAL row = b, used by the adaptive mode)
//   b(i)   solution weights
//   e(i)   error weights over S+1 stages: err = h * sum_i e(i) k_i   (5th minus embedded 4th order)
//
// Coefficients were checked against the order conditions up to order 5 (main) / 4 (embedded):
// see tests/test_tableaus.py, which parses this file.
#pragma once

namespace pcg {

#define PCG_CX __host__ __device__ static constexpr

struct Rk4 {
  static constexpr int S = 4;
  static constexpr bool EMBEDDED = false;
  PCG_CX double a(int i, int j) {
    constexpr double A[5][4] = {
        {0, 0, 0, 0},
        {0.5, 0, 0, 0},
        {0, 0.5, 0, 0},
        {0, 0, 1.0, 0},
        {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6}};
    return A[i][j];
  }
  PCG_CX double b(int i) { return a(4, i); }
  PCG_CX double e(int) { return 0.0; }
};

// Dormand & Prince 1980, RK5(4)7M  ("RK45")
struct Dopri5 {
  static constexpr int S = 6;
  static constexpr bool EMBEDDED = true;
  PCG_CX double a(int i, int j) {
    constexpr double A[7][6] = {
        {0, 0, 0, 0, 0, 0},
        {1.0 / 5, 0, 0, 0, 0, 0},
        {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
        {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
        {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
        {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
        {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
    return A[i][j];
  }
  PCG_CX double b(int i) { return a(6, i); }
  PCG_CX double e(int i) {
    constexpr double E[7] = {71.0 / 57600, 0, -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200, 22.0 / 525, -1.0 / 40};
    return E[i];
  }
};

// Tsitouras 2011, RK5(4)7: the pair behind diffrax.Tsit5 (reference integrator.py:57)
struct Tsit5 {
  static constexpr int S = 6;
  static constexpr bool EMBEDDED = true;
  PCG_CX double a(int i, int j) {
    constexpr double A[7][6] = {
        {0, 0, 0, 0, 0, 0},
        {0.161, 0, 0, 0, 0, 0},
        {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
        {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
        {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
        {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
        {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
    return A[i][j];
  }
  PCG_CX double b(int i) { return a(6, i); }
  PCG_CX double e(int i) {
    constexpr double E[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                             -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1.0 / 66};
    return E[i];
  }
};

// RODAS4 (Hairer & Wanner, Solving ODEs II, sec. IV.7 / rodas.f): six-stage stiffly accurate, L-stable Rosenbrock
// method of order 4 with an embedded third-order solution, in the transformed form that needs no matrix-vector
// product with the Jacobian:
//     W = I / (gamma h) - J(y0)
//     W U_i = f(y0 + sum_{j<i} a(i,j) U_j) + sum_{j<i} (c(i,j) / h) U_j,      i = 1..6
//     y1 = y0 + sum_{j<5} a(5,j) U_j + U_6          (a(5, .) = a(4, .) with a(5,4) = 1),     err = U_6
// The stiff counterpart of the explicit pairs: the reference integrates every model with CVODES' BDF
// (integrator.py:163-182), which follows a stiff transient (an ignited cstr, a fast extraction column) that no explicit
// pair can.  tests/test_tableaus.py checks the coefficients through the convergence order on a stiff nonlinear problem.
struct Rodas4 {
  static constexpr int S = 6;
  static constexpr bool EMBEDDED = true;
  static constexpr bool ROSENBROCK = true;
  static constexpr double gamma = 0.25;
  PCG_CX double a(int i, int j) {
    constexpr double A[7][6] = {
        {0, 0, 0, 0, 0, 0},
        {0.1544e1, 0, 0, 0, 0, 0},
        {0.9466785280815826, 0.2557011698983284, 0, 0, 0, 0},
        {0.3314825187068521e1, 0.2896124015972201e1, 0.9986419139977817, 0, 0, 0},
        {0.1221224509226641e1, 0.6019134481288629e1, 0.1253708332932087e2, -0.6878860361058950, 0, 0},
        {0.1221224509226641e1, 0.6019134481288629e1, 0.1253708332932087e2, -0.6878860361058950, 1.0, 0},
        {0, 0, 0, 0, 0, 0}};
    return A[i][j];
  }
  PCG_CX double c(int i, int j) {
    constexpr double C[6][6] = {
        {0, 0, 0, 0, 0, 0},
        {-0.56688e1, 0, 0, 0, 0, 0},
        {-0.2430093356833875e1, -0.2063599157091915, 0, 0, 0, 0},
        {-0.1073529058151375, -0.9594562251023355e1, -0.2047028614809616e2, 0, 0, 0},
        {0.7496443313967647e1, -0.1024680431464352e2, -0.3399990352819905e2, 0.1170890893206160e2, 0, 0},
        {0.8083246795921522e1, -0.7981132988064893e1, -0.3152159432874371e2, 0.1631930543123136e2, -0.6058818238834054e1, 0}};
    return C[i][j];
  }
  PCG_CX double b(int) { return 0.0; }      // the solution is formed from the stage values (see above)
  PCG_CX double e(int) { return 0.0; }
};

// explicit tableaus have no ROSENBROCK member
template <class TB, class = void>
struct is_rosenbrock { static constexpr bool value = false; };
template <class TB>
struct is_rosenbrock<TB, decltype(void(TB::ROSENBROCK))> { static constexpr bool value = true; };

}  // namespace pcg
