// invariants.cuh -- von Mises and principal values of a stress/strain tensor on the device, with the
// reference's exact branch structure (calcVonMises, calcPrincipalVals, src/vpmStress/strainAndStressUtils.f90:
// 484-555 -> FFaTensorTransforms.C:33-67,229-296 -> FFa::cubicSolve, FFaMath.C:61-142).  Shared by the
// single-step full-result kernel (k2_full.cu) and the results-database record kernels (io_rdb.cu).
#pragma once
#include "common.cuh"

namespace fsr {

// FFa::cubicSolve (FFaMath.C:61-142): same case analysis, same tolerances.
__device__ inline int cubic_solve(double A, double B, double C, double D, double* X)
{
  const double epsilon = 1.0e-16;
  if (fabs(A) > epsilon) {
    const double epsmall = 1.0e-96;  // pow(epsilon, 6)
    double P = (C - B * B / (3.0 * A)) / (3.0 * A);
    double Q = ((2.0 * B * B / (27.0 * A) - C / 3.0) * B / A + D) / (A + A);
    double W = Q * Q + P * P * P;
    if (W <= -epsmall && P < 0.0) {
      double FI = acos(-Q / sqrt(-P * P * P));
      X[0] = 2.0 * sqrt(-P) * cos(FI / 3.0);
      X[1] = -2.0 * sqrt(-P) * cos((FI + 3.14159265358979323846) / 3.0);
      X[2] = -2.0 * sqrt(-P) * cos((FI - 3.14159265358979323846) / 3.0);
    } else if (fabs(W) < epsmall && Q <= 0.0) {
      X[0] = 2.0 * pow(-Q, 1.0 / 3.0);
      X[1] = -0.5 * X[0];
      X[2] = X[1];
    } else if (W > -epsmall && Q + sqrt(W) <= 0.0 && Q - sqrt(W) <= 0.0) {
      X[0] = pow(-Q + sqrt(W), 1.0 / 3.0) + pow(-Q - sqrt(W), 1.0 / 3.0);
      X[1] = -0.5 * X[0];
      X[2] = X[1];
    } else if (W >= epsmall && fabs(Q) > epsmall && P > 0.0) {
      double FI = atan(sqrt(P * P * P) / fabs(Q));
      double KI = atan(copysign(pow(tan(0.5 * FI), 1.0 / 3.0), Q));
      X[0] = -2.0 * sqrt(P) / tan(KI + KI);
      X[1] = -0.5 * X[0];
      X[2] = X[1];
    } else
      return -3;
    W = B / (3.0 * A);
    X[0] -= W; X[1] -= W; X[2] -= W;
    return 3;
  } else if (fabs(B) > epsilon) {
    const double epsmall = 1.0e-64;  // pow(epsilon, 4)
    double P = C * C - 4.0 * B * D;
    if (P > 0.0) {
      double Q = sqrt(P);
      X[0] = (-C + Q) / (B + B);
      X[1] = (-C - Q) / (B + B);
    } else if (P > -epsmall) {
      X[0] = -C / (B + B);
      X[1] = X[0];
    } else
      return -2;
    return 2;
  } else if (fabs(C) > epsilon) {
    X[0] = -D / C;
    return 1;
  }
  return 0;
}

// principalValues (FFaTensorTransforms.C:229-286).  On failure P keeps its previous content,
// like the reference (the Fortran caller then reads stale values).
__device__ inline void principal_values(int ncmp, const double* S, double* P)
{
  if (ncmp == 3) {
    double Cq = -(S[0] + S[1]);
    double Dq = S[0] * S[1] - S[2] * S[2];
    double X[3];
    if (cubic_solve(0.0, 1.0, Cq, Dq, X) != 2) return;
    if (X[0] < X[1]) { double t = X[0]; X[0] = X[1]; X[1] = t; }
    P[0] = X[0]; P[1] = X[1];
  } else if (ncmp == 6) {
    double s11 = S[0], s22 = S[1], s33 = S[2], s12 = S[3], s13 = S[4], s23 = S[5];
    double B = -(s11 + s22 + s33);
    double C = s11 * s22 + s11 * s33 + s22 * s33 - s12 * s12 - s13 * s13 - s23 * s23;
    double D = s11 * s23 * s23 + s22 * s13 * s13 + s33 * s12 * s12 - s11 * s22 * s33 - 2.0 * s12 * s13 * s23;
    double X[3];
    if (cubic_solve(1.0, B, C, D, X) != 3) return;
    double t;
    if (X[0] < X[1]) { t = X[0]; X[0] = X[1]; X[1] = t; }
    if (X[1] < X[2]) { t = X[1]; X[1] = X[2]; X[2] = t; }
    if (X[0] < X[1]) { t = X[0]; X[0] = X[1]; X[1] = t; }
    P[0] = X[0]; P[1] = X[1]; P[2] = X[2];
  } else
    P[0] = S[0];
}

__device__ inline double von_mises(int ncmp, const double* S)
{
  if (ncmp == 3) return sqrt(S[0] * S[0] + S[1] * S[1] - S[0] * S[1] + 3.0 * S[2] * S[2]);
  if (ncmp == 6)
    return sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2] - S[0] * S[1] - S[1] * S[2] - S[2] * S[0] +
                3.0 * (S[3] * S[3] + S[4] * S[4] + S[5] * S[5]));
  return S[0];
}

}  // namespace fsr
