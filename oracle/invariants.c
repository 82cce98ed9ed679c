/* invariants.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): von Mises, principal values.
 * Follows fedem-foundation/src/FFaLib/FFaAlgebra/FFaTensorTransforms.C:33-67,229-296,335-361,
 * FFaMath.C:61-142 (FFa::cubicSolve), src/vpmStress/strainAndStressUtils.f90:14-98,484-555.
 * Validated bit-for-bit against the reference's own compiled C++ (oracle/_ref) in tests/. */
#include "oracle.h"
#include <math.h>
#include <stddef.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* FFaTensorTransforms.C:33-67 */
double orc_von_mises(int N, const double *S)
{
  switch (N) {
  case 1:
    return S[0];
  case 2:
    return sqrt(S[0] * S[0] + S[1] * S[1] - S[0] * S[1] + 3.0 * S[2] * S[2]);
  case 3:
    return sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2] - S[0] * S[1] - S[1] * S[2] -
                S[2] * S[0] + 3.0 * (S[3] * S[3] + S[4] * S[4] + S[5] * S[5]));
  default:
    return -HUGE_VAL;
  }
}

/* FFaMath.C:61-142 */
int orc_cubic_solve(double A, double B, double C, double D, double *X)
{
  const double epsilon = 1.0e-16;

  if (fabs(A) > epsilon) {
    const double epsmall = pow(epsilon, 6.0);
    double P = (C - B * B / (3.0 * A)) / (3.0 * A);
    double Q = ((2.0 * B * B / (27.0 * A) - C / 3.0) * B / A + D) / (A + A);
    double W = Q * Q + P * P * P;

    if (W <= -epsmall && P < 0.0) {
      double FI = acos(-Q / sqrt(-P * P * P));
      X[0] = 2.0 * sqrt(-P) * cos(FI / 3.0);
      X[1] = -2.0 * sqrt(-P) * cos((FI + M_PI) / 3.0);
      X[2] = -2.0 * sqrt(-P) * cos((FI - M_PI) / 3.0);
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
      return -3; /* incl. "case gamma not implemented" */

    W = B / (3.0 * A);
    for (int i = 0; i < 3; i++) X[i] -= W;
    return 3;
  } else if (fabs(B) > epsilon) {
    const double epsmall = pow(epsilon, 4.0);
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

/* FFaTensorTransforms.C:229-286; returns 1 on success, 0 on failure (P left untouched
 * past what cubicSolve wrote, exactly like the reference). */
int orc_principal_values(int N, const double *S, double *P)
{
  double t;
  switch (N) {
  case 1:
    P[0] = S[0];
    return 1;
  case 2: {
    double C = -(S[0] + S[1]);
    double D = S[0] * S[1] - S[2] * S[2];
    if (orc_cubic_solve(0.0, 1.0, C, D, P) != 2) return 0;
    if (P[0] < P[1]) { t = P[0]; P[0] = P[1]; P[1] = t; }
    return 1;
  }
  case 3: {
    double s11 = S[0], s22 = S[1], s33 = S[2], s12 = S[3], s13 = S[4], s23 = S[5];
    double B = -(s11 + s22 + s33);
    double C = s11 * s22 + s11 * s33 + s22 * s33 - s12 * s12 - s13 * s13 - s23 * s23;
    double D = s11 * s23 * s23 + s22 * s13 * s13 + s33 * s12 * s12 - s11 * s22 * s33 -
               2.0 * s12 * s13 * s23;
    if (orc_cubic_solve(1.0, B, C, D, P) != 3) return 0;
    if (P[0] < P[1]) { t = P[0]; P[0] = P[1]; P[1] = t; }
    if (P[1] < P[2]) { t = P[1]; P[1] = P[2]; P[2] = t; }
    if (P[0] < P[1]) { t = P[0]; P[0] = P[1]; P[1] = t; }
    return 1;
  }
  }
  return 0;
}

/* FFaTensorTransforms.C:295 */
double orc_max_shear_value(double pmax, double pmin) { return 0.5 * (pmax - pmin); }

/* FFaTensorTransforms.C:335-361 (rotate2D: eX = rotMx[0:2], eY = rotMx[2:4]); in-place safe */
void orc_rotate2d(const double *S, const double *rotMx, double *out)
{
  const double *eX = rotMx, *eY = rotMx + 2;
  double TS11 = eX[0] * S[0] + eY[0] * S[2];
  double TS12 = eX[0] * S[2] + eY[0] * S[1];
  double TS21 = eX[1] * S[0] + eY[1] * S[2];
  double TS22 = eX[1] * S[2] + eY[1] * S[1];
  out[0] = TS11 * eX[0] + TS12 * eY[0];
  out[1] = TS21 * eX[1] + TS22 * eY[1];
  out[2] = TS11 * eX[1] + TS12 * eY[1];
}

static int ncalc_of(int ncomp)
{
  switch (ncomp) {
  case 1: return 1;
  case 3: return 2;
  case 6: return 3;
  }
  return 0;
}

/* strainAndStressUtils.f90:484-516 */
void orc_calc_von_mises(const double *inTensor, int ncomp, int nstrp, double *vm)
{
  int ncalc = ncalc_of(ncomp);
  for (int i = 0; i < nstrp; i++) vm[i] = orc_von_mises(ncalc, inTensor + (size_t)ncomp * i);
}

/* strainAndStressUtils.f90:519-555; prinValVec persists across points like the Fortran local */
void orc_calc_principal_vals(const double *inTensor, int ncomp, int nstrp, double *maxP,
                             double *minP, double *maxS)
{
  int ncalc = ncalc_of(ncomp);
  double pv[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < nstrp; i++) {
    orc_principal_values(ncalc, inTensor + (size_t)ncomp * i, pv);
    maxP[i] = pv[0];
    minP[i] = pv[ncalc - 1];
    maxS[i] = orc_max_shear_value(pv[0], pv[ncalc - 1]);
  }
}

/* strainAndStressUtils.f90:14-57 */
void orc_principle_strains2d(const double epsC[3], double *eps1, double *eps2,
                             double *gammaMax, double *alpha1, double *alphaGamma)
{
  double origo = (epsC[0] + epsC[1]) * 0.5;
  double eps_12 = epsC[0] - epsC[1];
  double eps_xy = epsC[2] * 0.5;
  double radius = sqrt(eps_12 * eps_12 + epsC[2] * epsC[2]) * 0.5;
  *eps1 = origo + radius;
  *eps2 = origo - radius;
  *gammaMax = radius * 2.0;
  if (fabs(eps_xy) > ORC_EPSDIV0 || fabs(eps_12) > ORC_EPSDIV0) {
    if (alpha1) *alpha1 = atan2(eps_xy, eps_12) * 0.5;
    if (alphaGamma) *alphaGamma = atan2(eps_12, eps_xy) * 0.5;
  } else {
    if (alpha1) *alpha1 = 0.0;
    if (alphaGamma) *alphaGamma = 0.0;
  }
}

/* strainAndStressUtils.f90:60-98 */
void orc_principle_stresses2d(const double sigC[3], double *sig1, double *sig2,
                              double *tauMax, double *alpha1, double *alphaTau)
{
  double origo = (sigC[0] + sigC[1]) * 0.5;
  double sig_12 = sigC[0] - sigC[1];
  double radius = sqrt(sig_12 * sig_12 + 4.0 * sigC[2] * sigC[2]) * 0.5;
  *sig1 = origo + radius;
  *sig2 = origo - radius;
  *tauMax = radius;
  if (fabs(sigC[2]) > ORC_EPSDIV0 || fabs(sig_12) > ORC_EPSDIV0) {
    if (alpha1) *alpha1 = atan2(sigC[2], sig_12) * 0.5;
    if (alphaTau) *alphaTau = atan2(sig_12, sigC[2]) * 0.5;
  } else {
    if (alpha1) *alpha1 = 0.0;
    if (alphaTau) *alphaTau = 0.0;
  }
}
