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

}  // namespace pcg
